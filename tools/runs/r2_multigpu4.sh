#!/bin/bash
# round 2, 4 x B200: SURVEY 8(d) config 5 (Qwen2.5-VL video shape) and config 3 at N = 4
set +e
O=gpurun_out/r2mg4
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port"
timeout 900 $T 29521 bench.py --gpus 4 --model qwen2.5-vl-7b --video --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_video_n4.json 2> $O/bench_qwen_video_n4.err; echo "rc=$?"; cut -c1-2000 $O/bench_qwen_video_n4.json
timeout 900 $T 29522 bench.py --gpus 4 --model qwen2.5-vl-7b --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_n4.json 2> $O/bench_qwen_n4.err; echo "rc=$?"; cut -c1-1200 $O/bench_qwen_n4.json
echo done
