#!/bin/bash
# Round-1 run 26: re-validation after the RMSNorm dispatch change (suite, smoke, both bench lines).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run26.log
rm -f $L gpurun_out/check.jsonl
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke rc=$?" >> $L
timeout 100 python tools/gpu_check.py norm_perf 43008 4096 >> $L 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_r26.json 2>> $L; echo "bench rc=$?" >> $L
timeout 400 python bench.py --model qwen2.5-vl-7b > gpurun_out/bench_qwen_r26.json 2>> $L
python - <<PY >> $L 2>&1
import json
for f in ("bench_r26", "bench_qwen_r26"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"])
PY
grep -vE "^$|Warning|warn" $L | cut -c1-300 | tail -14
