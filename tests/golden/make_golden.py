"""Generates the committed golden vectors from the CPU oracle (run here, where the oracle builds):

    python tests/golden/make_golden.py

 * registration_case.npz — a self-contained L3 case (SURVEY §8b inner boundary): map points, keypoints, initial
   frame, previous frame; expected GN normal equations, GN poses, CERES poses, neighbor lists of 64 queries.
 * sequence_small16.json — per-frame outputs of Odometry::RegisterFrame on the seeded SMALL16 sequence (inputs are
   regenerated from the seed; a checksum of the inputs is stored to tell an input drift from an algorithm drift).

The reference itself holds no golden vectors for this path and cannot be built offline: these vectors
pin the ORACLE (regression guard) and give the GPU tests a fixed target that does not depend on rebuilding it.
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ct_icp_b200 import _abi as abi  # noqa: E402
from ct_icp_b200 import synthetic as syn  # noqa: E402
from oracle_lib import oracle  # noqa: E402


def frame_to_list(f):
    return [list(f.begin_pose.quat), list(f.begin_pose.tr), list(f.end_pose.quat), list(f.end_pose.tr),
            [f.begin_pose.dest_timestamp, f.end_pose.dest_timestamp]]


def frame_from_list(v, fid=1):
    f = abi.Frame()
    f.begin_pose = abi.Pose.make(v[0], v[1], v[4][0], fid)
    f.end_pose = abi.Pose.make(v[2], v[3], v[4][1], fid)
    return f


def sequence_options(b, solver):
    o = b.default_odometry_options()
    o.ct_icp_options.solver = abi.SOLVER[solver]
    o.ct_icp_options.min_number_neighbors = 10
    o.ct_icp_options.ls_max_num_iters = 5
    o.ct_icp_options.ls_num_threads = 1
    o.map_options = b.legacy_map_options(1.0, 20, 0.1)
    o.init_num_frames = 4
    o.debug_print = 0
    return o


def inputs_digest(seq):
    h = hashlib.sha256()
    for s in seq:
        h.update(np.ascontiguousarray(s["xyz"]).tobytes())
        h.update(np.ascontiguousarray(s["t"]).tobytes())
    return h.hexdigest()


def main():
    orc = oracle()
    # ---- L3 registration case ---------------------------------------------------------------------------------
    seq = syn.make_sequence(4, syn.SMALL16, seed=1234)
    od = orc.odometry(sequence_options(orc, "GN"))
    for s in seq[:3]:
        assert od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]).success
    map_xyz, _ = od.GetMapPointer().export(0)
    s = seq[3]
    idx = orc.grid_sample_indices(s["xyz"], 1.0)
    kp = np.zeros(len(idx), dtype=abi.wpoint_dtype())
    kp["raw"] = s["xyz"][idx]
    kp["timestamp"] = s["t"][idx]
    traj = od.Trajectory()
    prev = traj[-1]
    frame = traj[-1].copy()
    frame.begin_pose = traj[-1].end_pose.copy()
    frame.begin_pose.dest_timestamp = float(s["t"].min())
    frame.end_pose.dest_timestamp = float(s["t"].max())
    out = (C.c_double * 3)()
    for i in range(len(kp)):
        raw = (C.c_double * 3)(*kp["raw"][i])
        orc.check(orc.fn("pose_transform")(C.byref(frame), raw, float(kp["timestamp"][i]), out))
        kp["world"][i] = out[:]
    m = orc.voxel_map(orc.legacy_map_options(1.0, 20, 0.1))
    m.insert(map_xyz)
    mm = orc.default_odometry_options().default_motion_model
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    io.min_number_neighbors = 10
    A, b, n_used = m.gn_normal_equations(io, kp, frame, prev, mm)
    f_gn = frame.copy()
    s_gn = m.icp_register(io, kp.copy(), f_gn, prev, mm)
    io2 = orc.default_icp_options()
    io2.solver = abi.SOLVER["CERES"]
    io2.min_number_neighbors = 10
    io2.ls_max_num_iters = 5
    io2.ls_num_threads = 1
    io2.max_num_residuals = 400
    f_ce = frame.copy()
    s_ce = m.icp_register(io2, kp.copy(), f_ce, prev, mm, abi.StrategyOptions(0, 20, 8, 0))
    assert s_gn.success and s_ce.success
    queries = kp["world"][:64].copy()
    nb, cnt = m.compute_neighborhoods(queries, 20)
    np.savez_compressed(
        os.path.join(HERE, "registration_case.npz"),
        map_xyz=map_xyz.astype(np.float64), kp_raw=kp["raw"], kp_t=kp["timestamp"], kp_world=kp["world"],
        frame=np.array(sum(frame_to_list(frame)[:4], [])), frame_ts=np.array(frame_to_list(frame)[4]),
        prev=np.array(sum(frame_to_list(prev)[:4], [])), prev_ts=np.array(frame_to_list(prev)[4]),
        gn_A=A, gn_b=b, gn_n_used=np.array([n_used]),
        gn_frame=np.array(sum(frame_to_list(f_gn)[:4], [])), gn_residuals=np.array([s_gn.num_residuals_used]),
        ceres_frame=np.array(sum(frame_to_list(f_ce)[:4], [])), ceres_residuals=np.array([s_ce.num_residuals_used]),
        nb_queries=queries, nb_points=nb, nb_counts=cnt)
    print("registration_case.npz: map %d pts, %d keypoints, GN used %d, CERES used %d" %
          (len(map_xyz), len(kp), s_gn.num_residuals_used, s_ce.num_residuals_used))

    # ---- odometry sequence --------------------------------------------------------------------------------------
    seq = syn.make_sequence(8, syn.SMALL16, seed=1234)
    out = {"sensor": "SMALL16", "seed": 1234, "frames": 8, "inputs_sha256": inputs_digest(seq), "solvers": {}}
    for solver in ("GN", "CERES"):
        od = orc.odometry(sequence_options(orc, solver))
        rows = []
        for s in seq:
            sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
            rows.append({"success": int(sm.success), "F": int(sm.num_corrected_points), "K": int(sm.num_keypoints),
                         "residuals": int(sm.number_of_residuals), "map_size": int(od.MapSize()),
                         "pose": frame_to_list(sm.frame)})
        out["solvers"][solver] = rows
    with open(os.path.join(HERE, "sequence_small16.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("sequence_small16.json written")


if __name__ == "__main__":
    main()
