#!/bin/bash
# One GPU-box visit, in priority order (each step under its own timeout; logs under gpurun_out/):
#   1. the whole GPU test-suite (all failures reported) + a bisection over the round-2 switches if anything fails
#   2. the bench line (and one without the fused kernels if the tests failed)
#   3. the profiling evidence (launch list + one --set full capture of the dominant kernel)         [all only]
#   4. the A/B of the experiment builds (tools/ab_variants.sh)                                      [all only]
#   5. compute-sanitizer memcheck / racecheck / synccheck                                           [all only]
#   6. the bench line with the extra workloads and the CPU baselines                                [all only]
# usage (on the box): bash tools/gpu_check.sh <tag> [pytest|bench|quick|all]
TAG=${1:-x}; WHAT=${2:-all}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
RC=0
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
if [ "$WHAT" != "bench" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
  RC=$?
  echo "rc=$RC" >> gpurun_out/${TAG}_pytest.log
  tail -40 gpurun_out/${TAG}_pytest.log
  python tools/summarize_parity.py gpurun_out gpurun_out/${TAG}_parity_worst.json > gpurun_out/${TAG}_parity_worst.txt 2>&1
  if [ $RC -ne 0 ]; then
    # which switch breaks it? a fast subset under each fallback
    SUB="sequence_small or gn_register or gn_normal or ceres_register_matches_oracle or robust_register or neighborhoods"
    for cfg in "CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_selv1.so" "CTICP_FUSED_SAMPLING=0" "CTICP_FUSED_MAP_UPDATE=0" "CTICP_PERSISTENT=0" \
               "CTICP_FUSED_SAMPLING=0 CTICP_FUSED_MAP_UPDATE=0" "CTICP_FUSED_SAMPLING=0 CTICP_FUSED_MAP_UPDATE=0 CTICP_PERSISTENT=0"; do
      name=$(echo "$cfg" | tr ' =/' '___' | sed 's/.*libcticp_b200_//')
      env $cfg timeout 600 python -m pytest tests -m gpu -q -n 6 --tb=line -p no:cacheprovider -k "$SUB" > gpurun_out/${TAG}_bisect_${name}.log 2>&1
      echo "== $cfg: $(tail -1 gpurun_out/${TAG}_bisect_${name}.log)"
    done
  fi
fi
if [ "$WHAT" != "pytest" ]; then
  timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
  if [ $RC -ne 0 ]; then
    CTICP_FUSED_SAMPLING=0 CTICP_FUSED_MAP_UPDATE=0 timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_nofuse.json 2> gpurun_out/${TAG}_bench_nofuse.err
    echo "bench (no fused kernels) rc=$?"; tail -3 gpurun_out/${TAG}_bench_nofuse.err; cat gpurun_out/${TAG}_bench_nofuse.json
  fi
fi
if [ "$WHAT" = "all" ] || [ "$WHAT" = "knobs" ]; then
  # run-time knobs, one bench line each (step / GN loop / e2e): CTAs per SM of the two fused kernels (grid-barrier cost vs
  # latency hiding), keypoints per gather CTA, and the loop without its solve (CTICP_DEBUG_FLAGS=1: timing only)
  echo "---- knobs"
  for cfg in "CTICP_SAMPLE_CTAS_PER_SM=1" "CTICP_SAMPLE_CTAS_PER_SM=2" "CTICP_UPDATE_CTAS_PER_SM=1" "CTICP_UPDATE_CTAS_PER_SM=2" \
             "CTICP_GN_KP_PER_CTA=16" "CTICP_DEBUG_FLAGS=1"; do
    name=$(echo "$cfg" | tr ' =' '__')
    env $cfg timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_knob_${name}.json 2> gpurun_out/${TAG}_knob_${name}.err
    python - "$cfg" gpurun_out/${TAG}_knob_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-32s step %.4f ms  GN loop %6.1f us  e2e %.4f ms  solve share %s" % (sys.argv[1], d["ms_per_step"],
          d["roofline"]["us_per_launch"], d["e2e"]["ms_per_step"], d["roofline"].get("serial_reduce_and_solve_share")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
  done
  # where the solver CTA's serial part goes (instrumented build; SM cycles per frame on stderr)
  CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 \
      2>&1 | grep "GN loop, solver CTA" | tail -4 | tee gpurun_out/${TAG}_solver_cta_stamps.log
  CTICP_ENGINE_LIB=$PWD/ct_icp_b200/libcticp_b200_timers.so CTICP_DEBUG_TIMERS=1 timeout 300 python tools/profile_step.py --frames 30 \
      --workload kitti64_ceres 2>&1 | grep "LM loop, solver CTA" | tail -6 | tee -a gpurun_out/${TAG}_solver_cta_stamps.log
fi
if [ "$WHAT" = "all" ]; then
  echo "---- profile"; bash tools/gpu_profile.sh ${TAG} kitti64_gn 2>&1 | tail -14
  # dense workload (K ~ 32k: 14 keypoints per warp tile): 9 frames, the capture is the GN launch of the last one
  echo "---- profile dense128"; bash tools/gpu_profile.sh ${TAG}_dense dense128_gn 9 7 2>&1 | tail -14
  echo "---- A/B"; timeout 2400 bash tools/ab_variants.sh run gpurun_out/${TAG}_ab 2>&1 | tail -12
  echo "---- sanitizer"; timeout 2400 bash tools/gpu_sanitize.sh ${TAG} 2>&1 | tail -12
  echo "---- full bench"; timeout 1500 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
  echo "full bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_full.err; cat gpurun_out/${TAG}_bench_full.json
  echo "---- reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  cat gpurun_out/${TAG}_bench_reference.json | cut -c1-600
fi
