"""Loads the CPU oracle (oracle/liboracle.so) through the generic cticp binding. TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

from ct_icp_b200._binding import Binding

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_cached = None


def build_oracle():
    subprocess.run(["make", "-C", _ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    return os.path.join(_ORACLE_DIR, "liboracle.so")


def oracle():
    global _cached
    if _cached is None:
        path = os.path.join(_ORACLE_DIR, "liboracle.so")
        srcs = [os.path.join(_ORACLE_DIR, f) for f in os.listdir(_ORACLE_DIR) if f.endswith((".h", ".cpp"))]
        srcs.append(os.path.join(_ROOT, "include", "cticp.h"))
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            try:
                build_oracle()
            except Exception:
                if not os.path.exists(path):
                    raise
        _cached = Binding(ctypes.CDLL(path), "orc_")
    return _cached
