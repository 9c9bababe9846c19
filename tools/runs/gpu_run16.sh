#!/bin/bash
# Round-1 run 16: full validation of the final build + the profile set committed under profiles/ (launch list, ncu --set full of the
# top kernels).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run16.log
rm -f $L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke rc=$?" >> $L
timeout 900 python bench.py > gpurun_out/bench_final.json 2>> $L; echo "bench rc=$?" >> $L
timeout 300 python tools/gpu_check.py preprocess_perf 64 512 336 >> $L 2>&1
echo "== launch list" >> $L
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_bf16|attn_tc|attn_kernel|rmsnorm|layernorm|patchify|clip_embed|splice|decoder_embed|bias_table|lse_finalize|transpose_bsd|cross_softmax|t5_decoder|t5_cross" \
   -s 702 -c 702 --csv --log-file gpurun_out/r01_launches.csv python bench.py --ncu >> $L 2>&1; echo "launch list rc=$?" >> $L
echo "== ncu full" >> $L
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -s 560 -c 4 -o gpurun_out/r01_gemm_full python bench.py --ncu >> $L 2>&1; echo "ncu gemm rc=$?" >> $L
timeout 900 ncu --set full --import-source on --clock-control none -k regex:attn_tc_d64 -s 40 -c 1 -o gpurun_out/r01_attn_full python bench.py --ncu >> $L 2>&1; echo "ncu attn rc=$?" >> $L
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gemm_bf16_sm100_kernel<128, 1, 4>" -s 1 -c 1 -o gpurun_out/r01_lmhead_full python bench.py --ncu >> $L 2>&1; echo "ncu lmhead rc=$?" >> $L
timeout 900 ncu --set full --import-source on --clock-control none -k regex:t5_rmsnorm -s 60 -c 1 -o gpurun_out/r01_rmsnorm_full python bench.py --ncu >> $L 2>&1; echo "ncu rmsnorm rc=$?" >> $L
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_preprocess -s 3 -c 1 -o gpurun_out/r01_preprocess_full python tools/gpu_check.py preprocess_perf 64 512 336 >> $L 2>&1; echo "ncu preprocess rc=$?" >> $L
ls -la gpurun_out/*.ncu-rep >> $L
grep -vE "^==PROF|^==WARN|^$|Warning|warn" $L | cut -c1-1500 | tail -30
