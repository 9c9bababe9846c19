#!/bin/bash
# Round-1 run 12: device pre-processing (parity + perf), bench with the device pre-processing inside e2e.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run12.log
rm -f $L
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -15 >> $L
timeout 300 python tools/gpu_check.py preprocess_perf 64 512 336 >> $L 2>&1
timeout 300 python tools/gpu_check.py preprocess_perf 64 1024 336 >> $L 2>&1
timeout 300 python tools/gpu_check.py preprocess_perf 16 3000 336 >> $L 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r12.json 2>> $L; echo "bench rc=$?" >> $L
python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_r12.json"))
print("bench", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"], d["e2e"])
PY
grep -v "Warning\|warn" $L | cut -c1-700 | tail -30
