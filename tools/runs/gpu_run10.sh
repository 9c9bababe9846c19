#!/bin/bash
# 2-GPU validation of the torchrun path (NCCL all-gather of scores, max-over-ranks timing) + reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run10.log
rm -f $L
nvidia-smi --query-gpu=index,name --format=csv >> $L 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2>> $L; echo "bench n2 rc=$?" >> $L
cat gpurun_out/bench_n2.json >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2>> $L; echo "bench ref n2 rc=$?" >> $L
cat gpurun_out/bench_ref_n2.json >> $L
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2>> $L; echo "bench n1 rc=$?" >> $L
cat gpurun_out/bench_n1.json >> $L
grep -vE "^$|Warning|warn" $L | cut -c1-1500 | tail -20
