#!/bin/bash
# Third pass: tcgen05 attention + absorbed cross-attention bring-up, full test suite, bench, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run3.log
rm -f $L gpurun_out/check.jsonl
run() { timeout 240 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
VQA_ATTN_MMA_SYNC=1 run $C attention
run $C attention_perf 64 672 64 1
run $C attention_perf 64 577 16 0
VQA_ATTN_MMA_SYNC=1 run $C attention_perf 64 672 64 1
run $C pipeline tiny 3
VQA_CROSS=reference run $C pipeline tiny 3
VQA_ATTN_MMA_SYNC=1 run $C pipeline tiny 3
run $C pipeline mid 4
VQA_CROSS=reference run $C pipeline mid 4
echo "== pytest gpu" >> $L
timeout 1500 python -m pytest tests/ -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
grep -E "^\[|vs oracle|engine \[|passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head -60 >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "smoke rc=$?" >> $L
echo "== bench" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r1b.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1b.json >> $L
K='regex:gemm_bf16|attn_tc|flash_attn|t5_|layernorm|clip_embed|patchify|splice|decoder_embed|bias_table|lse_finalize|cross_softmax|transpose_bsd'
echo "== ncu launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 900 -c 900 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --ncu >> $L 2>&1; echo "ncu list rc=$?" >> $L
echo "== ncu full (GEMM + attention)" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 -s 500 -c 4 -o gpurun_out/prof_gemm_r1 \
    python bench.py --ncu >> $L 2>&1; echo "ncu full gemm rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 30 -c 2 -o gpurun_out/prof_attn_r1 \
    python bench.py --ncu >> $L 2>&1; echo "ncu full attn rc=$?" >> $L
tail -50 $L
