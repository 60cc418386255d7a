#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 900 python tools/debug_used_mismatch.py 24 all > gpurun_out/${TAG}_mismatch.log 2>&1; tail -26 gpurun_out/${TAG}_mismatch.log
echo "---- stamps"
CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 2>&1 | grep "GN loop, solver CTA" | tail -3 | tee gpurun_out/${TAG}_solver_cta_stamps.log
CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 --workload kitti64_ceres 2>&1 | grep "LM loop, solver CTA" | tail -4 | tee -a gpurun_out/${TAG}_solver_cta_stamps.log
echo "---- selv1 / prefetch2"
for v in selv1 prefetch2; do CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_$v.so timeout 600 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'step %.4f GN loop %.1f e2e %.4f'%(d['ms_per_step'], d['roofline']['us_per_launch'], d['e2e']['ms_per_step']))"; done
echo "---- profile"; bash tools/gpu_profile.sh ${TAG} kitti64_gn 2>&1 | tail -6
