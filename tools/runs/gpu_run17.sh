#!/bin/bash
# Round-1 run 17: attention d64 softmax restructure (single TMEM wait, far-tile constant bias, folded exponent FFMA).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run17.log
rm -f $L gpurun_out/check.jsonl
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4 >> $L
for i in 1 2; do
VQA_BIAS_CONST=1 timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
VQA_BIAS_CONST=0 timeout 300 python tools/gpu_check.py attention_perf 64 672 64 1 >> $L 2>&1
done
timeout 300 python tools/gpu_check.py attention_perf 64 577 16 0 >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_clipt5.py tests/test_gpu_qwen.py -m gpu -x -q 2>&1 | tail -4 >> $L
timeout 600 python bench.py --model qwen2.5-vl-7b > gpurun_out/bench_qwen_r17.json 2>> $L; python -c "import json; d=json.load(open(\"gpurun_out/bench_qwen_r17.json\")); print(\"qwen\", round(d[\"value\"],2), d[\"breakdown_ms\"])" >> $L 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r17.json 2>> $L; echo "bench rc=$?" >> $L
python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_r17.json"))
print("bench", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"], d["e2e"]["value"])
PY
grep -vE "^$|Warning|warn" $L | cut -c1-400 | tail -20
