#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run9.log
rm -f $L gpurun_out/check.jsonl
run() { timeout 300 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
if grep -q '"ok": false' gpurun_out/check.jsonl || ! grep -q '"test": "attention"' gpurun_out/check.jsonl; then echo "EARLY FAILURE" >> $L; cat gpurun_out/check.jsonl >> $L; tail -20 $L; exit 1; fi
run $C attention_perf 64 672 64 1
run $C attention_perf 64 577 16 0
echo "== pytest gpu" >> $L
timeout 1200 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
tail -3 gpurun_out/pytest_gpu.log >> $L
echo "== bench clipt5" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1h.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1h.json >> $L
echo "== bench qwen" >> $L
timeout 900 python bench.py --model qwen2.5-vl-7b --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_qwen_r1b.json 2>> $L; echo "bench qwen rc=$?" >> $L
cat gpurun_out/bench_qwen_r1b.json >> $L
echo "== ncu gemm traffic (all GEMM launches of step 2)" >> $L
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 -s 432 -c 432 --csv \
   --log-file gpurun_out/gemm_traffic_r1.csv python bench.py --ncu >> $L 2>&1; echo "ncu traffic rc=$?" >> $L
grep -vE "^==PROF|^==WARN|^$" $L | cut -c1-900 | tail -30
