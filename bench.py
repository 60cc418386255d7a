#!/usr/bin/env python
"""bench.py — RegisterFrame throughput of the B200-native CT-ICP engine on BASELINE.json's metric.

A "step" is one cticp RegisterFrame of one synthetic KITTI-shape 64-beam scan (~130k points) with the driving
options of BASELINE.json configs[1] (solver GN, 5 ICP iterations, map voxel 1.0 m / 20 pts, voxel_size 0.5,
sample_voxel_size 1.5). Steps are consecutive frames of ONE odometry run: the first `--preroll` frames (the
reference's start-up regime, init_num_frames = 20) and the W warm-up frames are registered untimed.

  value     scans/s with the packed scans already resident in HBM (cticp_odometry_register_staged), timed per step
            with CUDA events on the engine's stream, L2 flushed (untimed 256 MiB memset) between steps
  e2e       scans/s through cticp_odometry_register_frame with HOST numpy buffers: host packing into pinned
            memory, H2D of the scan, all kernels, D2H of poses/counters, wall clock per step incl. the map-update tail
  roofline  neighbor-gather kernel (k_gn_gather): algorithmic bytes per launch / CUDA-event time per launch vs the
            measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restatement of the reference's GN path: serial per-keypoint loop like
            src/ct_icp/ct_icp.cpp:753) on the same frames, on this box's host cores

`--impl reference` times only the CPU oracle (the reference itself cannot be built offline).
N > 1 (torchrun): every rank registers the same scans with the keypoints sharded rank/world and one NCCL
all-reduce of the 12x12 normal equations per iteration ("strong" scaling of one frame's latency).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "scans/sec (RegisterFrame) on 64-beam ~120k-pt clouds"
UNIT = "scans/s"
WORKLOAD = "configs[1]: KITTI-shape 64-beam synthetic scans, CT_ICP_GN point-to-plane, 5 ICP iters, 1xB200"


WORKLOADS = {
    # name: (sensor, description)
    "kitti64_gn": ("HDL64", WORKLOAD),
    "kitti64_ceres": ("HDL64", "configs[2]: KITTI-shape 64-beam synthetic scans, driving_config.yaml (solver CERES as device "
                               "LM/IRLS, Cauchy, 5x5 iterations, 900 residuals), 1xB200"),
    "dense128_gn": ("DENSE128", "configs[4]: dense 128-beam ~290k-pt synthetic scans, CT_ICP_GN, 20 ICP iterations forced, "
                                "voxel 0.25 / sample 0.5 (K ~ 19k keypoints)"),
}
_WORKLOAD = "kitti64_gn"


def make_options(b):
    from ct_icp_b200 import _abi as abi
    if _WORKLOAD == "kitti64_ceres":
        o = b.profile("default_driving")
        o.debug_print = 0
        o.neighborhood_strategy.max_num_neighbors = 20
        o.neighborhood_strategy.min_num_neighbors = 10
        m = b.default_map_options()
        m.num_resolutions = 1
        m.resolutions[0].resolution = 0.8
        m.resolutions[0].max_num_points = 30
        m.resolutions[0].min_distance_between_points = 0.1
        m.default_radius = 0.75
        o.map_options = m
        c = o.ct_icp_options
        c.debug_print = 0
        c.num_iters_icp, c.solver, c.max_num_residuals = 5, abi.SOLVER["CERES"], 900
        c.min_number_neighbors = c.max_number_neighbors = 20
        c.threshold_orientation_norm, c.threshold_translation_norm = 0.1, 0.01
        c.loss_function, c.ls_max_num_iters, c.ls_num_threads, c.ls_sigma = abi.LOSS["CAUCHY"], 5, 6, 0.1
        return o
    if _WORKLOAD == "dense128_gn":
        o = b.default_odometry_options()
        o.debug_print = 0
        o.ct_icp_options.solver = abi.SOLVER["GN"]
        o.ct_icp_options.num_iters_icp = 20
        o.ct_icp_options.threshold_orientation_norm = 0.0
        o.ct_icp_options.min_number_neighbors = 10
        o.map_options = b.legacy_map_options(1.0, 20, 0.1)
        o.voxel_size = o.init_voxel_size = 0.25
        o.sample_voxel_size = o.init_sample_voxel_size = 0.5
        return o
    o = b.default_odometry_options()
    o.ct_icp_options.solver = abi.SOLVER["GN"]
    o.ct_icp_options.num_iters_icp = 5
    o.ct_icp_options.min_number_neighbors = 10      # test/regression/regression_config_short_drive.yaml:93
    o.ct_icp_options.max_number_neighbors = 20
    o.ct_icp_options.max_dist_to_plane_ct_icp = 0.3
    o.map_options = b.legacy_map_options(1.0, 20, 0.1)   # size_voxel_map 1.0, 20 pts/voxel, min_distance 0.1
    o.voxel_size = 0.5
    o.sample_voxel_size = 1.5
    o.max_distance = 100.0
    o.debug_print = 0
    o.ct_icp_options.debug_print = 0
    return o


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    In-process NVML (pynvml) from a daemon thread every 20 ms (the timed regions are only tens of ms long; at 5 ms the queries began to show in the step times) — two cheap queries per sample, no child process next
    to the HOST-timed end-to-end steps; `nvidia-smi -lms 200` is the fallback (CTICP_BENCH_CLOCKS=smi forces it)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.sm, self.smax, self.reasons = [], [], set()
        self.mode = os.environ.get("CTICP_BENCH_CLOCKS", "nvml")
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                for name, mask in self.REASONS.items():
                    if bits & mask:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                self.sm.append(float(f[0]))
                self.smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def start(self):
        if self.mode == "nvml":
            try:
                import pynvml as nv
                nv.nvmlInit()
                # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = self.gpu_index
                if vis and all(v.strip().isdigit() for v in vis.split(",")):
                    idx = int(vis.split(",")[self.gpu_index])
                handle = nv.nvmlDeviceGetHandleByIndex(idx)
                self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.mode = "smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": max(self.smax) if self.smax else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "sampler": self.mode}


def run_oracle(seq, preroll, warmup, steps):
    """CPU oracle over the same frames; returns (scans/s over the timed steps, per-step ms list, counters)."""
    from oracle_lib import oracle
    orc = oracle()
    od = orc.odometry(make_options(orc))
    times = []
    for i, s in enumerate(seq[:preroll + warmup + steps]):
        t0 = time.perf_counter()
        sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        dt = time.perf_counter() - t0
        if not sm.success:
            raise RuntimeError("oracle registration failed at frame %d: %s" % (i, sm.error_message))
        if i >= preroll + warmup:
            times.append(dt * 1e3)
    return len(times) / (sum(times) / 1e3), times, od


def load_traffic():
    """dram bytes per launch of k_gn_gather from the committed ncu capture summary (profiles/), if any."""
    path = os.path.join(ROOT, "profiles", "gather_kernel_ncu_summary.json")
    try:
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--preroll", type=int, default=20, help="start-up frames registered untimed before the warm-up")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--roofline-frames", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="kitti64_gn", choices=sorted(WORKLOADS),
                    help="kitti64_gn is BASELINE.json's metric configuration (the bench line); the others are extra "
                         "measurements recorded under profiles/")
    args = ap.parse_args()
    global _WORKLOAD, WORKLOAD
    _WORKLOAD = args.workload
    sensor_name, WORKLOAD = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3)
    K = args.steps
    n_roof = args.roofline_frames if args.impl == "native" else 0

    from ct_icp_b200 import synthetic as syn
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        seq = syn.make_sequence(args.preroll + W + K, getattr(syn, sensor_name), seed=1234)
        npts = float(np.mean([len(s["xyz"]) for s in seq]))
        v, times, _ = run_oracle(seq, args.preroll, W, K)
        ms = float(np.mean(times))
        from oracle_lib import oracle as _orc
        ref_threads = int(make_options(_orc()).ct_icp_options.ls_num_threads)
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_scan": npts, "preroll_frames": args.preroll,
                       "note": "CPU restatement (oracle/) of the reference's RegisterFrame; the reference cannot be "
                               "built offline (Eigen/Ceres/glog/yaml-cpp/robin_map absent)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref_threads, "kind": "port",
                             "sample": "%d consecutive steady-state frames after %d untimed frames; threading as in the "
                                       "reference: GN per-keypoint loop serial (src/ct_icp/ct_icp.cpp:753), point "
                                       "transforms (and the CERES/ROBUST residual assembly) on ls_num_threads = %d OpenMP "
                                       "threads (odometry.cpp:469,480; ct_icp.cpp:561); %d host CPUs on this box"
                                       % (K, args.preroll + W, ref_threads, cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------ native arm
    import torch
    import ct_icp_b200
    eng = ct_icp_b200.engine()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if world > 1 else 0

    seq = syn.make_sequence(args.preroll + W + K + n_roof, getattr(syn, sensor_name), seed=1234)
    npts = float(np.mean([len(s["xyz"]) for s in seq]))
    n_timed_begin = args.preroll + W

    shard_modes = []

    def make_odometry():
        od = eng.odometry(make_options(eng), device)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                import ctypes
                buf = (ctypes.c_char * 128)()
                eng.check(eng.fn("nccl_unique_id")(buf))
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
            dist.broadcast(uid, 0)
            od.enable_sharding(bytes(uid.cpu().numpy().tobytes()), rank, world)
            shard_modes.append(od.sharding_mode())
        return od

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(device)

    # ---- pass A: device-resident input, CUDA-event timing per step ------------------------------------------
    od = make_odometry()
    slots = [od.stage_frame(s["xyz"], s["t"]) for s in seq]
    for i in range(n_timed_begin):
        sm = od.RegisterStaged(slots[i], seq[i]["frame_idx"])
        assert sm.success, sm.error_message
    od.last_timing()
    barrier()
    clocks.start()
    step_ms, launches, kp_sum, f_sum, iters_sum = [], 0, 0, 0, 0
    for i in range(n_timed_begin, n_timed_begin + K):
        od.flush_l2(256 << 20)
        if dist is not None:
            barrier()                     # all ranks receive the scan at the same time (untimed)
        od.timer_start()
        sm = od.RegisterStaged(slots[i], seq[i]["frame_idx"])
        ms = od.timer_stop()
        assert sm.success, sm.error_message
        step_ms.append(ms)
        t = od.last_timing()
        launches += t.kernel_launches
        kp_sum += sm.num_keypoints
        f_sum += sm.num_corrected_points
        iters_sum += t.icp_iterations
    barrier()
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    value = K / (total_ms / 1e3)

    # ---- roofline pass: per-launch CUDA events around k_gn_gather (continues pass A's odometry) ---------------
    od.set_gather_timing(True)
    g_ms, g_launch, g_kp, g_pts = 0.0, 0, 0, 0
    for i in range(n_timed_begin + K, n_timed_begin + K + n_roof):
        od.flush_l2(256 << 20)
        sm = od.RegisterStaged(slots[i], seq[i]["frame_idx"])
        t = od.last_timing()
        g_ms += t.gather_ms
        g_launch += t.gather_launches
        g_kp += t.gather_keypoint_iterations
        g_pts += t.gather_stencil_points
    od.set_gather_timing(False)
    stencil = 27        # (2r+1)^3 with r = ceil(0.8 / 1.0) = 1
    alg_bytes = g_kp * (16 + 16 * stencil) + 16 * g_pts          # keypoint + slot probes + map points
    peak, peak_src = measured_peak_gbs()
    if g_launch and g_ms > 0:
        achieved = (alg_bytes / g_launch) / (g_ms / g_launch * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_gn_persistent (all ICP iterations of a frame: gather + reduce + solve)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": load_traffic(), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / g_launch, "us_per_launch": g_ms / g_launch * 1e3,
                    "keypoints_per_launch": g_kp / g_launch, "mean_stencil_points": g_pts / max(g_kp, 1),
                    "launches_timed": g_launch}
    else:
        roofline = None
    od.clear_staged()
    od.close()

    # ---- pass B: end to end through the C ABI with host buffers ----------------------------------------------
    od = make_odometry()
    for i in range(n_timed_begin):
        sm = od.RegisterFrame(seq[i]["xyz"], seq[i]["t"], seq[i]["frame_idx"])
        assert sm.success, sm.error_message
    od.last_timing()
    barrier()
    e2e_ms, h2d, d2h = [], 0, 0
    for i in range(n_timed_begin, n_timed_begin + K):
        od.flush_l2(256 << 20)
        torch.cuda.synchronize(device)
        od.last_timing()                  # the previous frame's tail has completed
        if dist is not None:
            barrier()                     # all ranks receive the scan at the same time (untimed)
        t0 = time.perf_counter()
        sm = od.RegisterFrame(seq[i]["xyz"], seq[i]["t"], seq[i]["frame_idx"])
        t = od.last_timing()              # waits for the map-update tail of this frame
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
        assert sm.success, sm.error_message
        h2d += t.h2d_bytes
        d2h += t.d2h_bytes
    barrier()
    clock_info = clocks.stop()     # sampled from the start of the device-timed steps to the end of the e2e steps
    e2e_total = max_over_ranks(float(np.sum(e2e_ms)))
    e2e_value = K / (e2e_total / 1e3)
    od.close()

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, times, _ = run_oracle(seq, args.preroll, W, K)
        from oracle_lib import oracle as _orc
        ref_threads = int(make_options(_orc()).ct_icp_options.ls_num_threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": ref_threads, "kind": "port",
                        "sample": "the same %d timed frames (after %d untimed), CPU oracle restating the reference's "
                                  "RegisterFrame with the reference's threading: GN per-keypoint loop serial "
                                  "(src/ct_icp/ct_icp.cpp:753), point transforms on ls_num_threads = %d OpenMP threads "
                                  "(odometry.cpp:469,480); %d host CPUs on this box" % (K, n_timed_begin, ref_threads, cores),
                        "ms_per_step": float(np.mean(times))}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace("1xB200", "%dxB200" % world), "points_per_scan": npts, "frame_points": f_sum / K,
                       "keypoints": kp_sum / K, "icp_iters_per_step": iters_sum / K, "preroll_frames": args.preroll,
                       "l2": "flushed between steps (256 MiB memset, untimed)",
                       "parallelism": "single GPU" if world == 1 else "keypoints sharded x%d, JTJ/JTr summed over ranks once per iteration: %s" % (
                           world, "inside the persistent GN kernel over NVLink peer mailboxes (one launch per frame)"
                           if shard_modes and min(shard_modes) == 2 else "ncclAllReduce (peer mapping unavailable)"),
                       "storage": "fp32 voxel-local map points / keypoints, fp64 arithmetic"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_total / K, "ms_per_step_median": float(np.median(e2e_ms)),
                    "ms_per_step_max": float(np.max(e2e_ms)), "h2d_bytes_per_step": h2d / K,
                    "d2h_bytes_per_step": d2h / K, "timing": "wall clock per step incl. host packing and the map-update tail"},
            "gpu_launches": launches,
            "clocks": clock_info,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
