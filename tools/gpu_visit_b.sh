#!/bin/bash
# GPU visit: tests of the committed build, segment timings of the staged path (device tail on / off), A/B of the voxel prune,
# per-warp phase stamps of the GN gather (timers build). usage (on the box): bash tools/gpu_visit_b.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
timeout 1200 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -15 gpurun_out/${TAG}_pytest.log
python tools/summarize_parity.py gpurun_out gpurun_out/${TAG}_parity_worst.json > gpurun_out/${TAG}_parity_worst.txt 2>&1
echo "---- segments, device tail"; timeout 300 python tools/profile_step.py --frames 32 2>&1 | tail -6 | tee gpurun_out/${TAG}_segments_devtail.log
echo "---- segments, host tail"; CTICP_DEVICE_TAIL=0 timeout 300 python tools/profile_step.py --frames 32 2>&1 | tail -6 | tee gpurun_out/${TAG}_segments_hosttail.log
echo "---- bench (prune)"; timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json; python - gpurun_out/${TAG}_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("step %.4f ms  GN loop %6.1f us  e2e %.4f ms  dropin %.4f" % (d["ms_per_step"], d["roofline"]["us_per_launch"], d["e2e"]["ms_per_step"], d["e2e_dropin"]["ms_per_step"]))
PY
echo "---- bench (no prune)"; CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_noprune.so timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_noprune.json 2> gpurun_out/${TAG}_bench_noprune.err; python - gpurun_out/${TAG}_bench_noprune.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("step %.4f ms  GN loop %6.1f us  e2e %.4f ms  dropin %.4f" % (d["ms_per_step"], d["roofline"]["us_per_launch"], d["e2e"]["ms_per_step"], d["e2e_dropin"]["ms_per_step"]))
PY
echo "---- per-warp stamps"; CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 2>&1 | grep "GN gather warps\|GN loop, solver" | tail -12 | tee gpurun_out/${TAG}_warp_stamps.log
