#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run5.log
rm -f $L gpurun_out/check.jsonl
run() { timeout 200 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
if grep -q '"ok": false' gpurun_out/check.jsonl || ! grep -q '"test": "attention"' gpurun_out/check.jsonl; then echo "EARLY FAILURE" >> $L; cat gpurun_out/check.jsonl >> $L; tail -20 $L; exit 1; fi
run $C attention_perf 64 672 64 1
run $C attention_perf 64 577 16 0
run $C gemm 2562 1200 2048 512 gated_gelu 0 0
echo "== pytest gpu" >> $L
timeout 1200 python -m pytest tests/ -q -m gpu -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head -20 >> $L
echo "== bench" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1d.json >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 80 -c 1 -o gpurun_out/prof_attn_r1 \
    python bench.py --ncu >> $L 2>&1; echo "ncu full attn rc=$?" >> $L
tail -30 $L
