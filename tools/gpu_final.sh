#!/bin/bash
# Final evidence of a build, one GPU-box visit (logs under gpurun_out/<tag>_*): GPU test-suite, the full bench line (extra
# workloads + CPU baselines), the reference arm, the ncu launch list + one --set full capture of k_gn_persistent, and
# compute-sanitizer memcheck / racecheck / synccheck over short GN and CERES sequences.
# usage (on the box): bash tools/gpu_final.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.txt
timeout 1200 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
python tools/summarize_parity.py gpurun_out gpurun_out/${TAG}_parity_worst.json > gpurun_out/${TAG}_parity_worst.txt 2>&1
echo "---- full bench"; timeout 1500 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
echo "full bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_full.err; cut -c1-1800 gpurun_out/${TAG}_bench_full.json
echo "---- reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-400 gpurun_out/${TAG}_bench_reference.json
echo "---- segments"; timeout 300 python tools/profile_step.py --frames 32 2>&1 | tail -4 | tee gpurun_out/${TAG}_segments.log
timeout 300 python tools/profile_step.py --frames 30 --workload kitti64_ceres 2>&1 | tail -3 | tee gpurun_out/${TAG}_segments_ceres.log
echo "---- stamps"; CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 2>&1 | grep "GN gather warps\|GN loop, solver" | tail -6 > gpurun_out/${TAG}_warp_stamps.log; tail -2 gpurun_out/${TAG}_warp_stamps.log | cut -c1-400
echo "---- profile"; bash tools/gpu_profile.sh ${TAG} kitti64_gn 2>&1 | tail -8
echo "---- sanitizer"; timeout 1500 bash tools/gpu_sanitize.sh ${TAG} 2>&1 | tail -8
