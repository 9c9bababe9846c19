#!/bin/bash
# A/B: the attention kernel before (old) and after (new) the softmax restructure, same box, isolated and in-step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run23.log
rm -f $L gpurun_out/check.jsonl
for i in 1 2; do
echo "old" >> $L; VQA_B200_LIB=alt_build/libvqa_oldattn.so timeout 120 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
echo "new" >> $L; timeout 120 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
done
for v in old new old new; do
  if [ $v = old ]; then export VQA_B200_LIB=alt_build/libvqa_oldattn.so; else unset VQA_B200_LIB; fi
  timeout 400 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/bench_ab_$v.json 2>> $L
  python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_ab_$v.json"))
print("bench $v", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"])
PY
done
grep -vE "^$|Warning|warn" $L | cut -c1-300
