#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2l}
mkdir -p gpurun_out
echo "== default"; CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^2[89] |host min/max|pack region" | tail -8
echo "== 32 threads"; CTICP_HOST_THREADS=32 CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^29 |host min/max|pack region" | tail -4
