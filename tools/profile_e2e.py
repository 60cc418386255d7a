"""Host-side breakdown of the end-to-end RegisterFrame call (numpy buffers → C ABI)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, ct_icp_b200
from ct_icp_b200 import synthetic as syn
eng = ct_icp_b200.engine()
seq = syn.make_sequence(30, syn.HDL64, seed=1234)
od = eng.odometry(bench.make_options(eng))
for i, s in enumerate(seq):
    od.last_timing()
    t0 = time.perf_counter()
    sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
    t1 = time.perf_counter()
    t = od.last_timing()
    t2 = time.perf_counter()
    if i >= 24:
        print(i, "call %.3f ms (C side %.3f: init %.3f, try_register %.3f, map %.3f) + tail wait %.3f | device total %.3f ingest %.3f icp %.3f map %.3f" %
              ((t1 - t0) * 1e3, sm.odometry_total, sm.odometry_initialization, sm.odometry_try_register, sm.odometry_map_update,
               (t2 - t1) * 1e3, t.total_ms, t.ingest_ms, t.icp_ms, t.map_update_ms))
