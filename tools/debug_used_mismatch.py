"""Diagnosis of a residual-count mismatch between engine and oracle (GN): after N frames of the suburb sequence both maps
are identical; the keypoints of the next frame are then tested ONE AT A TIME through cticp_icp_gn_normal_equations on
both sides (n_used of a single linearisation = did this keypoint pass the neighbor-count and distance gates), and the
mismatching ones are described (neighbor sets, singular values of the covariance, distance to the plane)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ct_icp_b200  # noqa: E402
from ct_icp_b200 import _abi as abi  # noqa: E402
from ct_icp_b200 import synthetic as syn  # noqa: E402
from oracle_lib import oracle  # noqa: E402
from test_gpu_parity import _sequence_options  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
eng, orc = ct_icp_b200.engine(), oracle()
seq = syn.make_sequence(frames + 1, syn.HDL64E, seed=1234, scene=syn.UrbanScene(1234, profile="suburb"))


def opts(b):
    o = _sequence_options(b, "GN")
    o.ct_icp_options.num_iters_icp = 5
    o.voxel_size, o.sample_voxel_size, o.max_distance = 0.5, 1.5, 100.0
    return o


ods = {}
for name, b in (("orc", orc), ("eng", eng)):
    od = b.odometry(opts(b))
    for s in seq[:frames]:
        sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
    ods[name] = (od, sm)
    print(name, "frame", frames - 1, "residuals", sm.number_of_residuals, "keypoints", sm.num_keypoints, "map", od.MapSize())
frame = ods["orc"][1].frame                      # pose pair of the last registered frame ~ where the next scan is
s = seq[frames]
kp = np.zeros(len(s["xyz"][::30]), dtype=abi.wpoint_dtype())
kp["raw"] = s["xyz"][::30]
t0, t1 = s["t"].min(), s["t"].max()
kp["timestamp"] = s["t"][::30]
fr = abi.Frame()
for dst, src, ts in ((fr.begin_pose, frame.end_pose, t0), (fr.end_pose, frame.end_pose, t1)):
    for i in range(4):
        dst.quat[i] = src.quat[i]
    for i in range(3):
        dst.tr[i] = src.tr[i]
    dst.dest_timestamp = ts
from scipy.spatial.transform import Rotation  # noqa: E402
R = Rotation.from_quat(list(frame.end_pose.quat)).as_matrix()
T = np.array(list(frame.end_pose.tr))
kp["world"] = kp["raw"] @ R.T + T               # the reference's GN enters with the caller's world points
maps = {k: v[0].GetMapPointer() for k, v in ods.items()}
used = {}
for name, m in maps.items():
    u = np.zeros(len(kp), dtype=np.int32)
    io = opts(eng if name == "eng" else orc).ct_icp_options
    for i in range(len(kp)):
        one = kp[i:i + 1].copy()
        _, _, n = m.gn_normal_equations(io, one, fr)
        u[i] = n
    used[name] = u
    print(name, "used", int(u.sum()), "of", len(kp))
bad = np.flatnonzero(used["orc"] != used["eng"])
print("mismatching keypoints:", len(bad))
for i in bad[:12]:
    q = R @ kp["raw"][i] + T
    no, co = maps["orc"].compute_neighborhoods(q[None, :], 20)
    ne, ce = maps["eng"].compute_neighborhoods(q[None, :], 20)
    same = co[0] == ce[0] and np.abs(no[0, :co[0]] - ne[0, :ce[0]]).max() < 1e-6
    pts = no[0, :co[0]]
    sv = np.linalg.svd(np.cov(pts.T, bias=True), compute_uv=True)
    normal = sv[2][2]
    print("kp %d used orc %d eng %d | neighbors %d / %d same=%s | sv %s | |n.(p - p0)| %.6f" % (
        i, used["orc"][i], used["eng"][i], co[0], ce[0], same, np.array2string(sv[1], precision=6),
        abs(normal @ (q - pts[0]))))
