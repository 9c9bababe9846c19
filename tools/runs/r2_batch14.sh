#!/bin/bash
# round 2, GPU batch 14: head_dim-128 attention with eight softmax warps (two threads per query row) vs the four-warp build
set +e
O=gpurun_out/r2b14
mkdir -p $O
PREV=$PWD/t2v_metrics_b200/libvqa_b200_prev.so
timeout 900 python -m pytest tests/test_gpu_qwen.py -x -q > $O/qwen.log 2>&1; echo "rc=$?" >> $O/qwen.log; tail -3 $O/qwen.log
timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s -k qwen > $O/fullwidth_qwen.log 2>&1; echo "rc=$?" >> $O/fullwidth_qwen.log; grep -E "^\[qwen|passed|failed|rc=" $O/fullwidth_qwen.log | cut -c1-300
for tag in new prev new prev; do
  if [ $tag = prev ]; then export VQA_B200_LIB=$PREV; else unset VQA_B200_LIB; fi
  timeout 300 python tools/bench_kernels.py attn128 >> $O/attn128_$tag.jsonl 2>> $O/attn128.err
done
unset VQA_B200_LIB
for f in new prev; do echo $f; cat $O/attn128_$f.jsonl; done
for tag in new prev new prev; do
  if [ $tag = prev ]; then export VQA_B200_LIB=$PREV; else unset VQA_B200_LIB; fi
  timeout 900 python bench.py --model qwen2.5-vl-7b --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_${tag}_$RANDOM.json 2>> $O/bench.err
  timeout 900 python bench.py --model qwen2.5-vl-7b --video --video-size 336 --no-cpu-baseline --no-hf-baseline > $O/bench_qwen336_${tag}_$RANDOM.json 2>> $O/bench.err
done
unset VQA_B200_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b14/bench_*.json')):
    d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],2), 'pairs/s', round(d['ms_per_step'],2), 'ms', d['breakdown_ms'], d['clocks']['sm_mhz'])
PY
timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -3 $O/gpu_suite.log | head -2
echo done
