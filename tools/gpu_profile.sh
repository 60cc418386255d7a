#!/bin/bash
# One GPU-box visit for the profiling evidence (profiles/): launch list of steady-state frames + one `--set full` capture
# of the dominant kernel. Numbers printed under ncu are never bench values. usage: bash tools/gpu_profile.sh <tag> [workload]
TAG=${1:-x}; WL=${2:-kitti64_gn}; FRAMES=${3:-24}; SKIP=${4:-21}
KERNEL=k_gn_persistent; [ "$WL" = "kitti64_ceres" ] && KERNEL=k_lm_persistent
mkdir -p gpurun_out
# launch list: per-launch device time of every kernel of the last frames (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/profile_step.py --frames $FRAMES --workload $WL > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/${TAG}_launches.log
# full capture of ONE steady-state launch of the dominant kernel (skip the start-up frames' launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 -f -o gpurun_out/${TAG}_${KERNEL} \
    python tools/profile_step.py --frames $FRAMES --workload $WL > gpurun_out/${TAG}_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/${TAG}_full.log
ncu -i gpurun_out/${TAG}_${KERNEL}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${KERNEL}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_${KERNEL}.ncu-rep --page source --csv > gpurun_out/${TAG}_${KERNEL}_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_${KERNEL}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_${KERNEL}_lines.csv 2>/dev/null
python tools/summarize_ncu.py lines gpurun_out/${TAG}_${KERNEL}_lines.csv gpurun_out/${TAG}_${KERNEL}_lines.json | head -10
ls -la gpurun_out/${TAG}_*
