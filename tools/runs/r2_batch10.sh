#!/bin/bash
# round 2, GPU batch 10: Qwen vision tower at the native head width (grouped qkv GEMM output, compact attention output) + two windows per attention tile
set +e
O=gpurun_out/r2b10
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_qwen.py -x -q > $O/qwen.log 2>&1; echo "rc=$?" >> $O/qwen.log; tail -4 $O/qwen.log
timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s -k qwen > $O/fullwidth_qwen.log 2>&1; echo "rc=$?" >> $O/fullwidth_qwen.log; grep -E "^\[qwen|passed|failed|rc=" $O/fullwidth_qwen.log | cut -c1-400
for v in pairs nopairs pairs nopairs; do
  if [ $v = nopairs ]; then export VQA_ATTN128_NO_PAIRS=1; else unset VQA_ATTN128_NO_PAIRS; fi
  timeout 900 python bench.py --model qwen2.5-vl-7b --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_${v}_$RANDOM.json 2>> $O/bench.err
done
unset VQA_ATTN128_NO_PAIRS
timeout 900 python bench.py --model qwen2.5-vl-7b --video --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_video.json 2>> $O/bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b10/bench_qwen*.json')):
    d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms', d['breakdown_ms'], 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
PY
timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -3 $O/gpu_suite.log
echo done
