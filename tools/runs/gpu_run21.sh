#!/bin/bash
# Round-1 run 21: two-chain ping-pong attention, with a short canary first (a deadlocked kernel must not eat the GPU budget).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run21.log
rm -f $L gpurun_out/check.jsonl
VQA_ATTN_CHAINS=2 timeout 120 python tools/gpu_check.py attention >> $L 2>&1
rc=$?
if [ $rc -ne 0 ] || grep -q '"ok": false' gpurun_out/check.jsonl; then echo "CANARY FAILED rc=$rc" >> $L; cat $L | cut -c1-400 | tail -20; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention 2>&1 | tail -4 >> $L
if ! tail -4 $L | grep -q "passed"; then echo "PYTEST ATTENTION FAILED" >> $L; tail -20 $L | cut -c1-400; exit 1; fi
for c in 2 1 2 1; do
VQA_ATTN_CHAINS=$c timeout 120 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
done
VQA_ATTN_CHAINS=2 timeout 120 python tools/gpu_check.py attention_perf 64 577 16 0 >> $L 2>&1
VQA_ATTN_CHAINS=1 timeout 120 python tools/gpu_check.py attention_perf 64 577 16 0 >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_clipt5.py -m gpu -x -q 2>&1 | tail -3 >> $L
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_r21.json 2>> $L; echo "bench rc=$?" >> $L
python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_r21.json"))
print("bench", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"], d["e2e"]["value"])
PY
grep -vE "^$|Warning|warn" $L | cut -c1-300 | tail -24
