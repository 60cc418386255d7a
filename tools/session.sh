#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2n}
mkdir -p gpurun_out
echo "---- full bench"; timeout 1500 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
echo "full bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_full.err; cat gpurun_out/${TAG}_bench_full.json | cut -c1-1800
echo "---- reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cat gpurun_out/${TAG}_bench_reference.json | cut -c1-400
echo "---- e2e host breakdown"; CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^2[6789] |host min/max|pack region|H2D" | tail -16 > gpurun_out/${TAG}_e2e_host_breakdown.log; tail -4 gpurun_out/${TAG}_e2e_host_breakdown.log
for th in 8 24 32; do echo "== CTICP_HOST_THREADS=$th"; CTICP_HOST_THREADS=$th CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_e2e.py 2>&1 | grep -E "^29 |pack region" | tail -2; done
python tools/h2d_bandwidth.py 2>&1 | tail -3
