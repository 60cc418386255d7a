"""Multi-GPU parity check (run under torchrun, one rank per GPU): keypoint-sharded registration with one exchange
(sum over ranks of JTJ/JTr) per Gauss-Newton iteration (GN) / per LM evaluation (CERES, ROBUST) must reproduce the
single-GPU poses and be identical on every rank. The exchange runs inside the ICP kernels over NVLink peer mailboxes
(sharding_mode 2); CTICP_P2P=0 forces the ncclAllReduce fallback (mode 1). Also prints the mean steady-state
RegisterFrame latency (wall clock, max over ranks) of the single and the sharded run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ct_icp_b200  # noqa: E402
from ct_icp_b200 import synthetic as syn  # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
eng = ct_icp_b200.engine()
frames = int(os.environ.get("CTICP_CHECK_FRAMES", "24"))
seq = syn.make_sequence(frames, syn.HDL64, seed=1234)


def options_for(solver):
    from ct_icp_b200 import _abi as abi
    if solver == "GN":
        bench._WORKLOAD = "kitti64_gn"
        return bench.make_options(eng)
    bench._WORKLOAD = "kitti64_ceres"           # driving_config.yaml: 5 x 5 LM iterations, 900-residual prefix
    o = bench.make_options(eng)
    if solver == "ROBUST":                      # regression_robust_config_short_drive.yaml
        c = o.ct_icp_options
        c.solver = abi.SOLVER["ROBUST"]
        c.max_num_residuals, c.min_number_neighbors, c.ls_max_num_iters = 1000, 8, 8
        c.threshold_linearity, c.threshold_planarity, c.outlier_distance, c.use_barycenter = 0.9, 0.8, 0.8, 1
    return o


def run(sharded, solver):
    od = eng.odometry(options_for(solver), local_rank)
    if sharded:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (ctypes.c_char * 128)()
            eng.check(eng.fn("nccl_unique_id")(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        od.enable_sharding(uid.cpu().numpy().tobytes(), rank, world)
    mode = od.sharding_mode()
    poses, ms = [], []
    for i, s in enumerate(seq):
        if sharded:
            dist.barrier()          # ranks enter the frame together, like a driver feeding all GPUs the same scan
        od.last_timing()
        t0 = time.perf_counter()
        sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        od.last_timing()
        if i >= 21:
            ms.append((time.perf_counter() - t0) * 1e3)
        assert sm.success, sm.error_message
        poses.append(list(sm.frame.begin_pose.quat) + list(sm.frame.begin_pose.tr) + list(sm.frame.end_pose.quat) +
                     list(sm.frame.end_pose.tr))
    od.close()
    return np.array(poses), mode, (float(np.mean(ms)) if ms else float("nan"))


ok = True
for solver in os.environ.get("CTICP_CHECK_SOLVERS", "GN,CERES,ROBUST").split(","):
    single, _, ms_single = run(False, solver)
    sharded, mode, ms_sharded = run(True, solver)
    lat = torch.tensor([ms_single, ms_sharded], dtype=torch.float64, device="cuda")
    dist.all_reduce(lat, op=dist.ReduceOp.MAX)
    t = torch.from_numpy(sharded).cuda()
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    if rank == 0:
        across = max(float((g - gathered[0]).abs().max()) for g in gathered)
        vs_single = float(np.abs(sharded - single).max())
        print("MULTIGPU %s world=%d frames=%d sharding_mode=%d max|sharded - single|=%.3e max|rank_i - rank_0|=%.3e "
              "latency ms/frame: single %.3f sharded %.3f"
              % (solver, world, frames, mode, vs_single, across, float(lat[0]), float(lat[1])))
        ok = ok and across == 0.0 and vs_single < (1e-7 if solver == "GN" else 1e-6)
if rank == 0:
    assert ok, "sharded registration diverged"
    print("MULTIGPU OK")
dist.barrier()
dist.destroy_process_group()
