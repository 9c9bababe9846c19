#!/bin/bash
# Round-1 run 11: new Qwen paths (repetition penalty bitmap, video grids), full GPU suite with the rebuilt .so, new bench modes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run11.log
rm -f $L
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|Error|qwen penalty|qwen video|assert" | cut -c1-400 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke rc=$?" >> $L
timeout 600 python bench.py --model qwen2.5-vl-7b --video > gpurun_out/bench_qwen_video.json 2>> $L; echo "qwen video rc=$?" >> $L
timeout 600 python bench.py --model qwen2.5-vl-7b > gpurun_out/bench_qwen.json 2>> $L; echo "qwen rc=$?" >> $L
timeout 600 python bench.py --ragged --no-cpu-baseline > gpurun_out/bench_ragged.json 2>> $L; echo "ragged rc=$?" >> $L
timeout 600 python bench.py --pairs 1250 --steps 1 > gpurun_out/bench_job1250.json 2>> $L; echo "job rc=$?" >> $L
for f in bench_qwen_video bench_qwen bench_ragged bench_job1250; do python - <<PY >> $L 2>&1
import json
try:
    d = json.load(open("gpurun_out/$f.json"))
    print("$f", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms", d.get("breakdown_ms"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$f failed", e)
PY
done
tail -40 $L
