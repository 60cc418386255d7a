#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
echo "== default"; CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^2[89] |host min/max|pack region" | tail -6
for i in 1 2 3; do echo "== bench e2e (run $i)"; timeout 600 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step %.4f e2e %.4f median %.4f dropin %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_median'], d['e2e_dropin']['ms_per_step']))"; done
