#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run19.log
rm -f $L gpurun_out/check.jsonl
VQA_B200_LIB=build_exp/libvqa_exp9.so timeout 300 python tools/gpu_check.py attention_phases 64 672 64 >> $L 2>&1
VQA_ATTN_ONE_CTA=1 VQA_B200_LIB=build_exp/libvqa_exp9.so timeout 300 python tools/gpu_check.py attention_phases 64 672 64 >> $L 2>&1
VQA_B200_LIB=build_exp/libvqa_exp9.so timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
grep -vE "^$|Warning|warn" $L | cut -c1-600
