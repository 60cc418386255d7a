#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2f}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
timeout 900 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
python tools/summarize_parity.py gpurun_out gpurun_out/${TAG}_parity_worst.json | tail -70
echo "---- stamps"
CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 2>&1 | grep "GN loop, solver CTA" | tail -2 | tee gpurun_out/${TAG}_solver_cta_stamps.log
CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 --workload kitti64_ceres 2>&1 | grep "LM loop, solver CTA" | tail -4 | tee -a gpurun_out/${TAG}_solver_cta_stamps.log
echo "---- full bench"; timeout 1500 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
echo "full bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_full.err; cat gpurun_out/${TAG}_bench_full.json
