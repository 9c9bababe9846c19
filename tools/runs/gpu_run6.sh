#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run6.log
rm -f $L gpurun_out/check.jsonl
run() { timeout 300 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
run $C attention_perf 64 672 64 1
run $C attention_perf 64 577 16 0
run $C pytest tests/test_gpu_qwen.py -k attention
run $C pytest tests/test_gpu_qwen.py -k "engine or invariance"
run $C pytest tests/test_gpu_kernels.py tests/test_gpu_clipt5.py
echo "== bench" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1e.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1e.json >> $L
grep -E "rc=|passed|failed|Error|error|qwen|dlogp|oracle fp32|attention_perf|pairs/sec" $L | cut -c1-400 | tail -60
