#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run13.log
rm -f $L
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -15 >> $L
timeout 300 python tools/gpu_check.py preprocess_perf 64 512 336 >> $L 2>&1
timeout 300 python tools/gpu_check.py preprocess_perf 64 1024 336 >> $L 2>&1
timeout 300 python tools/gpu_check.py preprocess_perf 16 3000 336 >> $L 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q -k "bit_exact" 2>&1 | tail -8 >> $L
grep -v "Warning\|warn" $L | cut -c1-700 | tail -40
