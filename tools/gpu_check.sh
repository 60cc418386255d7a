#!/bin/bash
# One GPU-box visit: the whole GPU test-suite (all failures reported), then the bench lines. Logs under gpurun_out/.
# usage (on the box): bash tools/gpu_check.sh <tag> [pytest|bench|all]
TAG=${1:-x}; WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$WHAT" = "all" ] || [ "$WHAT" = "pytest" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -40 gpurun_out/${TAG}_pytest.log
fi
if [ "$WHAT" = "all" ] || [ "$WHAT" = "bench" ]; then
  timeout 900 python bench.py --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
fi
