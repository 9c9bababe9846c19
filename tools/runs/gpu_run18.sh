#!/bin/bash
# Round-1 run 18: where does the d64 attention kernel's time go? Ablations (results are WRONG by construction, timing only).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run18.log
rm -f $L gpurun_out/check.jsonl
echo "baseline" >> $L
timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
echo "one CTA per SM" >> $L
VQA_ATTN_ONE_CTA=1 timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
for e in 1 2 3 4; do
  echo "exp $e (1 no MUFU, 2 no bias/max pass, 3 no P store, 4 no S load)" >> $L
  VQA_B200_LIB=build_exp/libvqa_exp$e.so timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
done
echo "exp 1 + one CTA" >> $L
VQA_ATTN_ONE_CTA=1 VQA_B200_LIB=build_exp/libvqa_exp1.so timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
grep -vE "^$|Warning|warn" $L | cut -c1-200
