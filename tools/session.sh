#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2i}
mkdir -p gpurun_out
python tools/h2d_bandwidth.py 2>&1 | tail -4
lscpu | grep -E "Model name|Socket|NUMA node|Thread|Core" | head -8
nvidia-smi topo -m 2>/dev/null | head -6
for th in 2 4 8 16 32; do echo "== CTICP_HOST_THREADS=$th"; CTICP_HOST_THREADS=$th CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^29 |host min/max|H2D" | tail -3; done
echo "== default"; CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^2[789] |host min/max|H2D" | tail -6
