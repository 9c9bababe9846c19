#!/bin/bash
# round 2: compute-sanitizer memcheck over the kernels added this round (small test shapes)
set +e
O=gpurun_out/r2san
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention_edge or attention_matches or split_k or fused_rmsnorm" > $O/memcheck_kernels.log 2>&1; echo "kernels rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck_kernels.log | tail -5
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_qwen.py -x -q -k "window_pairs or grouped or prefix or trace" > $O/memcheck_qwen.log 2>&1; echo "qwen rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck_qwen.log | tail -5
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_clipt5.py -x -q -k "micro or graph" > $O/memcheck_clipt5.log 2>&1; echo "clipt5 rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck_clipt5.log | tail -5
echo done
