#!/bin/bash
# round 2, final validation of the shipped build: smoke, full GPU suite, the driver's bench invocations, Qwen lines
set +e
O=gpurun_out/r2final
mkdir -p $O
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?"; tail -1 $O/smoke.log | cut -c1-300
echo "== full GPU suite"; timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -3 $O/gpu_suite.log
echo "== reference arm"; timeout 1500 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"; cut -c1-400 $O/bench_reference.json
echo "== default bench"; timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; cut -c1-600 $O/bench_default.json
echo "== qwen"; timeout 1200 python bench.py --model qwen2.5-vl-7b > $O/bench_qwen.json 2> $O/bench_qwen.err; echo "rc=$?"; cut -c1-300 $O/bench_qwen.json
timeout 1200 python bench.py --model qwen2.5-vl-7b --video > $O/bench_qwen_video.json 2> $O/bench_qwen_video.err; echo "rc=$?"; cut -c1-300 $O/bench_qwen_video.json
python - <<'PY'
import json
for f in ('bench_default','bench_qwen','bench_qwen_video'):
    d=json.load(open('gpurun_out/r2final/'+f+'.json'))
    print(f, round(d['value'],1), round(d['ms_per_step'],1), d['breakdown_ms'], 'e2e', round(d['e2e']['value'],1), 'hf', round(d['hf_gpu_baseline']['value'],1), 'cpu', round(d['cpu_baseline']['value'],3), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['gpu_launches'])
PY
echo done
