#!/bin/bash
# A/B of experiment builds of the engine (compile-time variants of the kernels) against the default build.
#
#   tools/ab_variants.sh build            # here (no GPU needed): nvcc cross-compiles ct_icp_b200/libcticp_b200_<name>.so
#   tools/ab_variants.sh run [outdir]     # on the GPU box (inside ONE gpurun call): full GPU test-suite + bench per build
#
# Variants (name:flags). Every variant must pass the whole `pytest -m gpu` suite before its bench line counts.
#   handoff     -DCTICP_HANDOFF                       point-to-point hand-off instead of two grid barriers per GN iteration
#   reducemlp   -DCTICP_REDUCE_MLP                    all partial rows of a warp in flight before the first add
#   handoffmlp  -DCTICP_HANDOFF -DCTICP_REDUCE_MLP
#   prune       -DCTICP_PRUNE                         drop candidates beyond the current k-th distance
#   prefetch6/8 -DCTICP_PREFETCH=6 / 8                6 / 8 chunks (192 / 256 stencil points) in flight per load batch
#   warps8      -DCTICP_GATHER_WARPS=8                the previous CTA shape (control)
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
VARIANTS=("handoff:-DCTICP_HANDOFF" "reducemlp:-DCTICP_REDUCE_MLP" "handoffmlp:-DCTICP_HANDOFF -DCTICP_REDUCE_MLP"
          "prune:-DCTICP_PRUNE" "prefetch6:-DCTICP_PREFETCH=6" "prefetch8:-DCTICP_PREFETCH=8" "warps8:-DCTICP_GATHER_WARPS=8")
case "${1:-}" in
build)
    for v in "${VARIANTS[@]}"; do
        name="${v%%:*}"; flags="${v#*:}"
        make -C "$ROOT/ct_icp_b200/csrc" -j8 BUILD="build_$name" OUT="../libcticp_b200_$name.so" EXTRA="$flags" \
            > /dev/null || { echo "build of $name failed"; exit 1; }
        echo "built libcticp_b200_$name.so ($flags)"
    done ;;
run)
    out="${2:-$ROOT/gpurun_out/ab}"; mkdir -p "$out"
    cd "$ROOT"
    python bench.py --no-cpu-baseline > "$out/bench_default.json" 2> /dev/null
    for v in "${VARIANTS[@]}"; do
        name="${v%%:*}"
        lib="$ROOT/ct_icp_b200/libcticp_b200_$name.so"
        [ -f "$lib" ] || continue
        CTICP_ENGINE_LIB="$lib" python -m pytest tests -m gpu -q -n 6 > "$out/pytest_$name.log" 2>&1
        echo "rc=$?" >> "$out/pytest_$name.log"
        CTICP_ENGINE_LIB="$lib" python bench.py --no-cpu-baseline > "$out/bench_$name.json" 2> /dev/null
    done
    python - "$out" <<'PY'
import glob, json, os, sys
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        name = os.path.basename(f)[6:-5]
        log = os.path.join(sys.argv[1], "pytest_%s.log" % name)
        tests = open(log).read().strip().splitlines()[-2:] if os.path.exists(log) else ["(default build)"]
        print("%-12s step %.4f ms  GN loop %.1f us  e2e %.4f ms  | %s" % (
            name, d["ms_per_step"], d["roofline"]["us_per_launch"], d["e2e"]["ms_per_step"], " ".join(tests)))
    except Exception as e:
        print(f, "unreadable:", e)
PY
    ;;
*) echo "usage: $0 build | run [outdir]"; exit 2 ;;
esac
