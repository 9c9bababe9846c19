#!/bin/bash
# Round-1 run 25: warp-per-row RMSNorm (correctness canary, perf, whole suite, bench).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run25.log
rm -f $L gpurun_out/check.jsonl
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k norm 2>&1 | tail -3 >> $L
if ! tail -3 $L | grep -q "passed"; then echo "NORM TEST FAILED" >> $L; tail -20 $L | cut -c1-400; exit 1; fi
timeout 100 python tools/gpu_check.py norm_perf 43008 4096 >> $L 2>&1
timeout 100 python tools/gpu_check.py norm_perf 10240 3584 >> $L 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $L
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_r25.json 2>> $L; echo "bench rc=$?" >> $L
timeout 400 python bench.py --model qwen2.5-vl-7b > gpurun_out/bench_qwen_r25.json 2>> $L
python - <<PY >> $L 2>&1
import json
for f in ("bench_r25", "bench_qwen_r25"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"])
PY
grep -vE "^$|Warning|warn" $L | cut -c1-300 | tail -14
