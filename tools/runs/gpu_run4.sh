#!/bin/bash
# Fourth pass: 8-warp GEMM epilogue with prefetch, attention v2; tests, bench, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run4.log
rm -f $L gpurun_out/check.jsonl
run() { timeout 240 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
run $C gemm 2562 1000 776 1032 store 1 1
if grep -q '"ok": false' gpurun_out/check.jsonl || ! grep -q '"test": "attention"' gpurun_out/check.jsonl; then echo "EARLY FAILURE" >> $L; cat gpurun_out/check.jsonl >> $L; tail -20 $L; exit 1; fi
run $C attention_perf 64 672 64 1
run $C attention_perf 64 577 16 0
run $C gemm_perf 2562 36928 4096 1024 quick_gelu
run $C gemm_perf 2562 36928 3072 1024
run $C gemm_perf 2562 36928 1024 1024
run $C gemm_perf 2562 43008 4096 4096
run $C gemm_perf 2562 43008 20480 4096 gated_gelu
run $C gemm_perf 2562 43008 4096 10240
echo "== pytest gpu" >> $L
timeout 1500 python -m pytest tests/ -q -m gpu -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
grep -E "^\[|vs oracle|engine \[|passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head -60 >> $L
echo "== bench" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1c.json >> $L
K='regex:gemm_bf16|attn_tc|flash_attn|t5_|layernorm|clip_embed|patchify|splice|decoder_embed|bias_table|lse_finalize|cross_softmax|transpose_bsd'
echo "== ncu launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 702 -c 702 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --ncu >> $L 2>&1; echo "ncu list rc=$?" >> $L
echo "== ncu full (GEMM + attention)" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 -s 540 -c 6 -o gpurun_out/prof_gemm_r1 \
    python bench.py --ncu >> $L 2>&1; echo "ncu full gemm rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 80 -c 2 -o gpurun_out/prof_attn_r1 \
    python bench.py --ncu >> $L 2>&1; echo "ncu full attn rc=$?" >> $L
tail -40 $L
