#!/bin/bash
# round 2: weak-scaling points N = 2 and N = 4 with the final build (N = 1 and N = 8: r2_final.sh / r2_multigpu8.sh)
set +e
O=gpurun_out/r2mg24
mkdir -p $O
for N in 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "N=$N rc=$?"; cut -c1-260 $O/bench_n$N.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29539 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > $O/bench_reference_n4.json 2> $O/bench_reference_n4.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference_n4.json
echo done
