#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run8.log
rm -f $L gpurun_out/check.jsonl
run() { timeout 300 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
if grep -q '"ok": false' gpurun_out/check.jsonl || ! grep -q '"test": "attention"' gpurun_out/check.jsonl; then echo "EARLY FAILURE" >> $L; cat gpurun_out/check.jsonl >> $L; tail -20 $L; exit 1; fi
run $C attention_perf 64 672 64 1
run $C attention_perf 64 577 16 0
run $C gemm_perf 2562 43008 20480 4096 gated_gelu
run $C gemm_perf 2562 43008 12288 4096
echo "== pytest gpu" >> $L
timeout 1200 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
tail -3 gpurun_out/pytest_gpu.log >> $L
echo "== bench clipt5" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1g.json >> $L
grep -vE "^==PROF|^==WARN|^$" $L | cut -c1-700 | tail -30
