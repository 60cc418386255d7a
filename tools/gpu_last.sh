#!/bin/bash
# Last visit of the round: tests of the committed build (sampler pre-clear + verdict written by k_gn_persistent ON), the same
# with both knobs off if anything fails, A/B bench of the knobs, then the full bench line + reference arm as the final evidence.
TAG=${1:-x}; mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("step %.4f ms  GN loop %6.1f us  e2e %.4f ms  dropin %s  launches %s" % (d["ms_per_step"], r.get("us_per_launch", 0),
          d["e2e"]["ms_per_step"], (d.get("e2e_dropin") or {}).get("ms_per_step"), d.get("gpu_launches")))
except Exception as e:
    print("unreadable:", e)
PY
}
timeout 900 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
RC=$?; echo "rc=$RC" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
python tools/summarize_parity.py gpurun_out gpurun_out/${TAG}_parity_worst.json > gpurun_out/${TAG}_parity_worst.txt 2>&1
if [ $RC -ne 0 ]; then
  CTICP_SAMPLE_PRECLEAR=0 CTICP_TAIL_IN_KERNEL=0 timeout 900 python -m pytest tests -m gpu -q -n 6 --tb=line -p no:cacheprovider > gpurun_out/${TAG}_pytest_knobs_off.log 2>&1
  echo "knobs off: rc=$? $(tail -1 gpurun_out/${TAG}_pytest_knobs_off.log)"
  CTICP_SAMPLE_PRECLEAR=0 timeout 600 python -m pytest tests -m gpu -q -n 6 --tb=line -p no:cacheprovider -k "small or tail or sampl or suburb" > gpurun_out/${TAG}_pytest_preclear_off.log 2>&1
  echo "preclear off (subset): $(tail -1 gpurun_out/${TAG}_pytest_preclear_off.log)"
  CTICP_TAIL_IN_KERNEL=0 timeout 600 python -m pytest tests -m gpu -q -n 6 --tb=line -p no:cacheprovider -k "small or tail or sampl or suburb" > gpurun_out/${TAG}_pytest_tailkernel_off.log 2>&1
  echo "tail-in-kernel off (subset): $(tail -1 gpurun_out/${TAG}_pytest_tailkernel_off.log)"
fi
echo "---- bench (defaults)"; timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; line gpurun_out/${TAG}_bench.json
echo "---- bench (knobs off)"; CTICP_SAMPLE_PRECLEAR=0 CTICP_TAIL_IN_KERNEL=0 timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_knobs_off.json 2> gpurun_out/${TAG}_bench_knobs_off.err; line gpurun_out/${TAG}_bench_knobs_off.json
echo "---- full bench"; timeout 600 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; line gpurun_out/${TAG}_bench_full.json
echo "---- reference arm"; timeout 300 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
echo "---- segments"; timeout 200 python tools/profile_step.py --frames 30 2>&1 | tail -3 | tee gpurun_out/${TAG}_segments.log
