#!/bin/bash
# 2-GPU box visit: the whole GPU test-suite (the 2-GPU tests run here), the sharded-vs-single check for the three solvers with
# the in-kernel exchange and the NCCL fallback, and the bench line at N = 2 (configs[1] and the dense workload).
# usage (gpurun --gpus 2): bash tools/gpu_visit_2gpu.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ';'; echo
timeout 1500 python -m pytest tests -m gpu -q -n 6 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -15 gpurun_out/${TAG}_pytest.log
echo "---- sharded vs single (peer mailboxes)"; CTICP_CHECK_FRAMES=14 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -12 | tee gpurun_out/${TAG}_multigpu_check.txt
echo "---- sharded vs single (NCCL fallback)"; CTICP_P2P=0 CTICP_CHECK_FRAMES=10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/multigpu_check.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -8 | tee -a gpurun_out/${TAG}_multigpu_check.txt
echo "---- bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
echo "rc=$?"; tail -3 gpurun_out/${TAG}_bench_2gpu.err | cut -c1-300; python - gpurun_out/${TAG}_bench_2gpu.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("N=2 step %.4f ms e2e %.4f ms  sharded_vs_single %s  parallelism %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("sharded_vs_single"), d["arm"]["parallelism"][:80]))
    for k, v in (d.get("extra_workloads") or {}).items():
        print("  ", k, "step %.4f e2e %.4f" % (v["ms_per_step"], v["e2e"]["ms_per_step"]))
except Exception as e:
    print("unreadable:", e)
PY
