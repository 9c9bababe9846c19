#!/bin/bash
# round 2, 8 x B200: SURVEY 8(d) config 4 (a job of 10 000 pairs, strong scaling) and the weak-scaling step at N = 8
set +e
O=gpurun_out/r2mg8
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
nvidia-smi --query-gpu=index,name,power.limit --format=csv > $O/nvsmi.txt
timeout 1200 $T 29511 bench.py --gpus 8 --pairs 10000 --steps 2 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_pairs10000_n8.json 2> $O/bench_pairs10000_n8.err; echo "rc=$?"; cut -c1-2500 $O/bench_pairs10000_n8.json
timeout 900 $T 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?"; cut -c1-1500 $O/bench_n8.json
echo done
