#!/bin/bash
# First bring-up pass on the GPU box: kernel checks, tiny pipeline parity, a first GEMM perf number.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/check.jsonl gpurun_out/run1.log
run() { timeout 240 "$@" >> gpurun_out/run1.log 2>&1; echo "rc=$? :: $*" >> gpurun_out/run1.log; }
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
C="python tools/gpu_check.py"
run $C norm
run $C attention
run $C gemm 2561 128 256 64
run $C gemm 2561 512 512 512
run $C gemm 1281 512 512 512
run $C gemm 641 512 512 512
run $C gemm 321 512 512 512
run $C gemm 2562 256 256 64
run $C gemm 2562 512 512 512
run $C gemm 1282 512 512 512
run $C gemm 2561 1000 776 1032 store 1 1
run $C gemm 2562 1000 776 1032 store 1 1
run $C gemm 2562 3000 1024 2048 quick_gelu 1 0
run $C gemm 2562 3000 1024 640 gelu 1 0
run $C gemm 2562 1200 2048 512 gated_gelu 0 0
run $C gemm 2561 1200 2048 512 gated_gelu 0 0
run $C gemm 1281 100 2048 512 gated_gelu 0 0
run $C gemm 641 100 1024 512 gated_gelu 0 0
run $C gemm 321 100 1024 4096 store 0 1
run $C lmhead
VQA_GEMM_SIMT=1 run $C pipeline tiny 3
run $C pipeline tiny 3
VQA_GEMM_VARIANT=2561 run $C pipeline tiny 3
VQA_GEMM_VARIANT=2562 run $C pipeline tiny 3
run $C pipeline mid 4
run $C gemm_perf 2562 43008 4096 4096
run $C gemm_perf 2561 43008 4096 4096
run $C gemm_perf 2562 43008 20480 4096 gated_gelu
run $C gemm_perf 2562 43008 4096 10240
run $C gemm_perf 1282 43008 4096 4096
run $C gemm_perf 321 128 4096 4096
run $C gemm_perf 641 128 4096 4096
tail -5 gpurun_out/run1.log
cat gpurun_out/check.jsonl | cut -c1-400
