#!/bin/bash
# round 2, GPU batch 1: new attention kernel (correctness + A/B timing), full-width parity, GEMM tile-order sweep (time + DRAM bytes), bench
set +e
O=gpurun_out/r2b1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi.txt 2>&1
echo "== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > $O/kernels.log 2>&1; echo "rc=$?" >> $O/kernels.log; tail -3 $O/kernels.log
echo "== attention A/B"
for v in 1 10 12 13; do VQA_ATTN_VARIANT=$v timeout 300 python tools/bench_kernels.py attn >> $O/attn_ab.jsonl 2>> $O/attn_ab.err; done
cat $O/attn_ab.jsonl
echo "== clipt5 golden tests"; timeout 900 python -m pytest tests/test_gpu_clipt5.py -x -q -s > $O/clipt5.log 2>&1; echo "rc=$?" >> $O/clipt5.log; tail -4 $O/clipt5.log
echo "== full-width parity"; timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s > $O/fullwidth.log 2>&1; echo "rc=$?" >> $O/fullwidth.log; grep -E "^\[|passed|failed|Error|rc=" $O/fullwidth.log | tail -40
echo "== gemm schedules: time"; timeout 600 python tools/bench_kernels.py gemm-time > $O/gemm_time.json 2> $O/gemm_time.err; head -c 3000 $O/gemm_time.json
echo "== gemm schedules: dram bytes (ncu)"
python tools/bench_kernels.py gemm-list > $O/gemm_list.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 --csv \
    --log-file $O/gemm_ncu.csv python tools/bench_kernels.py gemm-ncu > $O/gemm_ncu.log 2>&1; echo "ncu rc=$?"
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cat $O/bench.json
echo done
