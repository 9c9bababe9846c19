#!/bin/bash
# Round-1 run 20: two-chain ping-pong attention (CHAINS=2) vs the two-CTAs-per-SM kernel (CHAINS=1).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run20.log
rm -f $L gpurun_out/check.jsonl
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention 2>&1 | tail -4 >> $L
for c in 2 1 2 1; do
VQA_ATTN_CHAINS=$c timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
done
VQA_ATTN_CHAINS=2 timeout 300 python tools/gpu_check.py attention_perf 64 577 16 0 >> $L 2>&1
VQA_ATTN_CHAINS=1 timeout 300 python tools/gpu_check.py attention_perf 64 577 16 0 >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_clipt5.py -m gpu -x -q 2>&1 | tail -3 >> $L
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r20.json 2>> $L; echo "bench rc=$?" >> $L
python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_r20.json"))
print("bench", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"], d["e2e"]["value"])
PY
grep -vE "^$|Warning|warn" $L | cut -c1-300 | tail -20
