#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2g}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
timeout 900 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider -k "ceres or robust or nclt or distance or motion or golden" > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
echo "---- stamps"
CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 --workload kitti64_ceres 2>&1 | grep "LM loop, solver CTA" | tail -4 | tee -a gpurun_out/${TAG}_solver_cta_stamps.log
echo "---- ceres bench"; timeout 900 python bench.py --workload kitti64_ceres --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ceres step %.4f ms e2e %.4f (median %.4f) launches %s iters %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_median'], d['gpu_launches'], d['config'].get('icp_iters_per_step')))"
echo "---- profile ceres"; bash tools/gpu_profile.sh ${TAG}c kitti64_ceres 2>&1 | tail -4
ncu -i gpurun_out/${TAG}c_k_lm_persistent.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}c_k_lm_persistent_lines.csv 2>/dev/null
python tools/summarize_ncu.py lines gpurun_out/${TAG}c_k_lm_persistent_lines.csv gpurun_out/${TAG}c_k_lm_persistent_lines.json | head -12
