#!/bin/bash
# round 2, GPU batch 6: split-row stage with split-phase exchange + deferred P-slot wait; full GPU suite; smoke; Qwen benches
set +e
O=gpurun_out/r2b6
mkdir -p $O
echo "== attention kernel tests with the split stage"
VQA_ATTN_VARIANT=40 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention and not rounding" > $O/kernels_split.log 2>&1; echo "rc=$?" >> $O/kernels_split.log; tail -2 $O/kernels_split.log
VQA_ATTN_VARIANT=41 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "rounding_points" > $O/kernels_split_round.log 2>&1; echo "rc=$?" >> $O/kernels_split_round.log; tail -2 $O/kernels_split_round.log
echo "== attention A/B (d64): 30 streaming, 40/41 split-row"
for v in 30 40 41 30 40; do VQA_ATTN_VARIANT=$v ATTN_ROUND=$((v % 10)) timeout 300 python tools/bench_kernels.py attn >> $O/attn_ab.jsonl 2>> $O/attn_ab.err; done
cat $O/attn_ab.jsonl
echo "== clipt5 goldens with the split stage"; VQA_ATTN_VARIANT=40 timeout 900 python -m pytest tests/test_gpu_clipt5.py -x -q > $O/clipt5_split.log 2>&1; echo "rc=$?" >> $O/clipt5_split.log; tail -3 $O/clipt5_split.log
echo "== bench: streaming vs split stage in the full step"
for v in 30 40 30 40; do VQA_ATTN_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_attn$v.json 2> $O/bench_attn$v.err
python -c "
import json
d=json.load(open('$O/bench_attn$v.json')); print('$v', round(d['value'],2), 'pairs/s', d['breakdown_ms'], d['clocks']['sm_mhz'], 'MHz', d['sample_scores'])"; done
echo "== ncu split stage"
VQA_ATTN_VARIANT=40 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc_d64 -s 2 -c 1 -o $O/attn_split python tools/bench_kernels.py attn-one > $O/ncu_attn.log 2>&1; echo "ncu rc=$?"
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
echo "== full GPU suite (default build)"; timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -5 $O/gpu_suite.log
timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s > $O/fullwidth.log 2>&1; echo "rc=$?" >> $O/fullwidth.log; grep -E "^\[|passed|failed|Error|rc=|spread" $O/fullwidth.log | cut -c1-400 | tail -50
timeout 600 python -m pytest tests/test_gpu_qwen.py -q -s -k prefix > $O/qwen_prefix.log 2>&1; grep -E "dlog|passed|failed" $O/qwen_prefix.log | cut -c1-300
echo "== qwen benches (config 3, config 5 shape)"
timeout 1200 python bench.py --model qwen2.5-vl-7b > $O/bench_qwen.json 2> $O/bench_qwen.err; echo "rc=$?"; cut -c1-2500 $O/bench_qwen.json
timeout 1200 python bench.py --model qwen2.5-vl-7b --video > $O/bench_qwen_video.json 2> $O/bench_qwen_video.err; echo "rc=$?"; cut -c1-2500 $O/bench_qwen_video.json
timeout 1200 python bench.py --impl reference --model qwen2.5-vl-7b --steps 3 --warmup 1 > $O/bench_qwen_reference.json 2> $O/bench_qwen_reference.err; echo "rc=$?"; cut -c1-1500 $O/bench_qwen_reference.json
echo done
