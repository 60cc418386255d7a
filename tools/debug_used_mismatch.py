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

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng, orc = ct_icp_b200.engine(), oracle()
seq = syn.make_sequence(frames + 1, syn.HDL64E, seed=1234, scene=syn.UrbanScene(1234, profile="suburb"))


def opts(b):
    o = _sequence_options(b, "GN")
    o.ct_icp_options.num_iters_icp = 5
    o.voxel_size, o.sample_voxel_size, o.max_distance = 0.5, 1.5, 100.0
    return o


# pass 1: both arms over the sequence; the first frame whose residual count differs
first_bad, sums = None, {}
ods = {name: b.odometry(opts(b)) for name, b in (("orc", orc), ("eng", eng))}
for k, s in enumerate(seq[:frames]):
    for name, od in ods.items():
        sums[name] = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
    so, se = sums["orc"], sums["eng"]
    from conftest import frame_diff  # noqa: E402
    dt, dr = frame_diff(so.frame, se.frame)
    print("frame %2d  residuals orc %d eng %d | keypoints %d %d | map %d %d | pose diff %.3e m %.3e rad" % (
        k, so.number_of_residuals, se.number_of_residuals, so.num_keypoints, se.num_keypoints, ods["orc"].MapSize(),
        ods["eng"].MapSize(), dt, dr))
    if so.number_of_residuals != se.number_of_residuals and first_bad is None:
        first_bad = k
        kps = {name: od.keypoints() for name, od in ods.items()}
        final = {name: sums[name].frame for name in ods}
        init = {name: sums[name].initial_frame.copy() for name in ods}
        if len(sys.argv) <= 2:   # a third argument: keep going over the whole sequence (pose differences after the first flip)
            break
if first_bad is None:
    print("no mismatch in %d frames" % frames)
    sys.exit(0)
if len(sys.argv) > 2:
    sys.exit(0)
print("keypoint records identical (raw, timestamp):", np.array_equal(kps["orc"]["raw"], kps["eng"]["raw"]),
      np.array_equal(kps["orc"]["timestamp"], kps["eng"]["timestamp"]))

# pass 2: fresh arms up to the frame before → the maps the bad frame was registered against
ods2 = {name: b.odometry(opts(b)) for name, b in (("orc", orc), ("eng", eng))}
for s in seq[:first_bad]:
    for od in ods2.values():
        od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
maps = {k: v.GetMapPointer() for k, v in ods2.items()}
kp = kps["orc"].copy()
fr = final["orc"]
used = {}
for name, m in maps.items():
    u = np.zeros(len(kp), dtype=np.int32)
    io = opts(eng if name == "eng" else orc).ct_icp_options
    for i in range(len(kp)):
        one = kp[i:i + 1].copy()
        _, _, n = m.gn_normal_equations(io, one, fr)
        u[i] = n
    used[name] = u
    print(name, "used", int(u.sum()), "of", len(kp), "(single linearisation at the oracle's final pose pair)")
bad = np.flatnonzero(used["orc"] != used["eng"])
print("mismatching keypoints:", len(bad))
for i in bad[:12]:
    q = kp["world"][i]
    no, co = maps["orc"].compute_neighborhoods(q[None, :], 20)
    ne, ce = maps["eng"].compute_neighborhoods(q[None, :], 20)
    same = co[0] == ce[0] and np.abs(no[0, :co[0]] - ne[0, :ce[0]]).max() < 1e-6
    pts = no[0, :co[0]]
    sv = np.linalg.svd(np.cov(pts.T, bias=True), compute_uv=True)
    normal = sv[2][2]
    print("kp %d used orc %d eng %d | neighbors %d / %d same=%s | sv %s | |n.(p - p0)| %.6f" % (
        i, used["orc"][i], used["eng"][i], co[0], ce[0], same, np.array2string(sv[1], precision=6),
        abs(normal @ (q - pts[0]))))

# pass 3: the bad frame's registration replayed through the L3 entry point with 1, 2, … iterations, on both arms, from
# the estimate RegisterFrame started from — where do the two iterations part?
o = opts(eng)
k = first_bad
traj = ods2["orc"].Trajectory()
prev = traj[-1] if traj else None
print("initial estimates (end tr):", list(init["orc"].end_pose.tr)[:3], list(init["eng"].end_pose.tr)[:3])
for n_it in range(1, 17):
    res = {}
    for name, b in (("orc", orc), ("eng", eng)):
        ob = opts(b)
        io = ob.ct_icp_options
        io.num_iters_icp = n_it
        kq = kps["orc"].copy()
        fr0 = init["orc"].copy()
        mm = ob.default_motion_model if ob.with_default_motion_model and prev is not None else None
        sm3 = maps[name].icp_register(io, kq, fr0, prev if mm is not None else None, mm)
        res[name] = (sm3.num_residuals_used, sm3.num_iters, np.array(list(fr0.end_pose.tr)), np.array(list(fr0.begin_pose.tr)))
    print("iters %2d: n_used orc %d eng %d | iters run %d %d | end tr diff %.3e begin tr diff %.3e" % (
        n_it, res["orc"][0], res["eng"][0], res["orc"][1], res["eng"][1],
        np.linalg.norm(res["orc"][2] - res["eng"][2]), np.linalg.norm(res["orc"][3] - res["eng"][3])))
