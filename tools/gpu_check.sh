#!/bin/bash
# One GPU-box visit: the whole GPU test-suite (all failures reported), a bisection over the round-2 switches if anything
# fails, then the bench lines. Logs under gpurun_out/.   usage (on the box): bash tools/gpu_check.sh <tag> [pytest|bench|all]
TAG=${1:-x}; WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
RC=0
if [ "$WHAT" = "all" ] || [ "$WHAT" = "pytest" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
  RC=$?
  echo "rc=$RC" >> gpurun_out/${TAG}_pytest.log
  tail -40 gpurun_out/${TAG}_pytest.log
  if [ $RC -ne 0 ]; then
    # which switch breaks it? a fast subset under each fallback
    SUB="sequence_small or gn_register or gn_normal or ceres_register_matches_oracle or robust_register or neighborhoods"
    for cfg in "CTICP_FUSED_SAMPLING=0" "CTICP_FUSED_MAP_UPDATE=0" "CTICP_PERSISTENT=0" \
               "CTICP_FUSED_SAMPLING=0 CTICP_FUSED_MAP_UPDATE=0" "CTICP_FUSED_SAMPLING=0 CTICP_FUSED_MAP_UPDATE=0 CTICP_PERSISTENT=0"; do
      name=$(echo "$cfg" | tr ' =' '__')
      env $cfg timeout 600 python -m pytest tests -m gpu -q -n 6 --tb=line -p no:cacheprovider -k "$SUB" > gpurun_out/${TAG}_bisect_${name}.log 2>&1
      echo "== $cfg: $(tail -1 gpurun_out/${TAG}_bisect_${name}.log)"
    done
  fi
fi
if [ "$WHAT" = "all" ] || [ "$WHAT" = "bench" ]; then
  timeout 900 python bench.py --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
  if [ $RC -ne 0 ]; then
    CTICP_FUSED_SAMPLING=0 CTICP_FUSED_MAP_UPDATE=0 timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_nofuse.json 2> gpurun_out/${TAG}_bench_nofuse.err
    echo "bench (no fused kernels) rc=$?"; tail -3 gpurun_out/${TAG}_bench_nofuse.err; cat gpurun_out/${TAG}_bench_nofuse.json
  fi
fi
