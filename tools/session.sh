#!/bin/bash
# scratch driver of one GPU visit (edited per visit; tools/gpu_check.sh is the maintained one)
TAG=${1:-r2h}
mkdir -p gpurun_out; rm -f gpurun_out/parity_worst.*.json
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "multigpu or second_device or ceres_register or hdl64_ceres" > gpurun_out/${TAG}_pytest_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_2gpu.log; tail -8 gpurun_out/${TAG}_pytest_2gpu.log
echo "---- multigpu_check (peer mailboxes)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_p2p.txt 2>&1; tail -14 gpurun_out/${TAG}_multigpu_check_p2p.txt
echo "---- multigpu_check (nccl fallback)"
CTICP_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_nccl.txt 2>&1; tail -8 gpurun_out/${TAG}_multigpu_check_nccl.txt
echo "---- bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; tail -3 gpurun_out/${TAG}_bench_2gpu.err; cat gpurun_out/${TAG}_bench_2gpu.json | cut -c1-3000
