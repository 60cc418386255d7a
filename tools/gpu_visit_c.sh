#!/bin/bash
# GPU visit: tests of the committed build, A/B against the build kept as libcticp_b200_gridsync.so (configs[1] and the dense
# workload), per-warp phase stamps (timers build), cooperative-launch micro-benchmark. usage: bash tools/gpu_visit_c.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("step %.4f ms  GN loop %6.1f us (%.2f us / 1k kp-iters, frac %.4f)  e2e %.4f ms  dropin %s  K %.0f" % (
        d["ms_per_step"], r.get("us_per_launch", 0), r.get("us_per_1k_keypoint_iterations", 0), r.get("frac", 0),
        d["e2e"]["ms_per_step"], (d.get("e2e_dropin") or {}).get("ms_per_step"), d["config"]["keypoints"]))
except Exception as e:
    print("unreadable:", e)
PY
}
timeout 1200 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -15 gpurun_out/${TAG}_pytest.log
python tools/summarize_parity.py gpurun_out gpurun_out/${TAG}_parity_worst.json > gpurun_out/${TAG}_parity_worst.txt 2>&1
echo "---- bench"; timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; line gpurun_out/${TAG}_bench.json
if [ -f ct_icp_b200/libcticp_b200_gridsync.so ]; then
  echo "---- bench (grid barriers)"; CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_gridsync.so timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_static.json 2> gpurun_out/${TAG}_bench_static.err; line gpurun_out/${TAG}_bench_static.json
fi
echo "---- dense128"; timeout 600 python bench.py --workload dense128_gn --no-extras --no-cpu-baseline --steps 6 > gpurun_out/${TAG}_bench_dense.json 2> gpurun_out/${TAG}_bench_dense.err; line gpurun_out/${TAG}_bench_dense.json
if [ -f ct_icp_b200/libcticp_b200_gridsync.so ]; then
  echo "---- dense128 (grid barriers)"; CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_gridsync.so timeout 600 python bench.py --workload dense128_gn --no-extras --no-cpu-baseline --steps 6 > gpurun_out/${TAG}_bench_dense_static.json 2> gpurun_out/${TAG}_bench_dense_static.err; line gpurun_out/${TAG}_bench_dense_static.json
fi
echo "---- per-warp stamps"; CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 2>&1 | grep "GN gather warps\|GN loop, solver" | tail -6 | tee gpurun_out/${TAG}_warp_stamps.log

