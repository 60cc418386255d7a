#!/bin/bash
# compute-sanitizer evidence for profiles/: memcheck, racecheck (shared-memory hazards) and synccheck (barrier misuse)
# over a short GN and a short CERES sequence. usage (on the box): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for solver in GN CERES; do
    frames=4; [ "$tool" = "racecheck" ] && frames=3
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
        python tools/sanitize_target.py $solver $frames > gpurun_out/${TAG}_sanitizer_${tool}_${solver}.log 2>&1
    echo "$tool $solver rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target OK' gpurun_out/${TAG}_sanitizer_${tool}_${solver}.log | tr '\n' ' ')"
  done
done
