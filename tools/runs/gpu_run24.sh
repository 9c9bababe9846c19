#!/bin/bash
# A/B: old attention kernel vs old + constant far-tile bias (in-tree).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run24.log
rm -f $L gpurun_out/check.jsonl
timeout 120 python tools/gpu_check.py attention >> $L 2>&1
rc=$?
if [ $rc -ne 0 ] || grep -q '"ok": false' gpurun_out/check.jsonl; then echo "CANARY FAILED rc=$rc" >> $L; cat $L | cut -c1-400 | tail -20; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention 2>&1 | tail -3 >> $L
for i in 1 2; do
echo "old" >> $L; VQA_B200_LIB=alt_build/libvqa_oldattn.so timeout 120 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
echo "new" >> $L; timeout 120 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
done
for v in old new; do
  if [ $v = old ]; then export VQA_B200_LIB=alt_build/libvqa_oldattn.so; else unset VQA_B200_LIB; fi
  timeout 400 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/bench_ab_$v.json 2>> $L
  python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_ab_$v.json"))
print("bench $v", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"])
PY
done
grep -vE "^$|Warning|warn" $L | cut -c1-300 | tail -16
