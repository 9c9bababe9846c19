#!/bin/bash
# round 2, GPU batch 4: split-row softmax stage (8 softmax warps per CTA) A/B against the streaming stage; new GPU tests (CUDA graph, top-k trace,
# KV-prefix sharing); full default bench with CPU + HF baselines; reference arm; GEMM DRAM traffic over every launch of one step.
set +e
O=gpurun_out/r2b4
mkdir -p $O
echo "== kernel tests: default stage, split stage (40 fp32 scores / 41 reference rounding)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > $O/kernels.log 2>&1; echo "rc=$?" >> $O/kernels.log; tail -2 $O/kernels.log
VQA_ATTN_VARIANT=40 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention and not rounding" > $O/kernels_split.log 2>&1; echo "rc=$?" >> $O/kernels_split.log; tail -2 $O/kernels_split.log
VQA_ATTN_VARIANT=41 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "attention" > $O/kernels_split_round.log 2>&1; echo "rc=$?" >> $O/kernels_split_round.log; tail -2 $O/kernels_split_round.log
echo "== attention A/B (d64): 30/31 streaming, 40/41 split-row"
for v in 30 40 31 41 30 40; do VQA_ATTN_VARIANT=$v ATTN_ROUND=$((v % 10)) timeout 300 python tools/bench_kernels.py attn >> $O/attn_ab.jsonl 2>> $O/attn_ab.err; done
cat $O/attn_ab.jsonl
echo "== clipt5 goldens + CUDA graph test with the split stage"; VQA_ATTN_VARIANT=40 timeout 900 python -m pytest tests/test_gpu_clipt5.py -x -q > $O/clipt5_split.log 2>&1; echo "rc=$?" >> $O/clipt5_split.log; tail -3 $O/clipt5_split.log
echo "== qwen GPU tests (top-k trace, prefix sharing)"; timeout 900 python -m pytest tests/test_gpu_qwen.py -x -q > $O/qwen.log 2>&1; echo "rc=$?" >> $O/qwen.log; tail -3 $O/qwen.log
echo "== bench: streaming vs split stage in the full step"
for v in 30 40; do VQA_ATTN_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_attn$v.json 2> $O/bench_attn$v.err
python -c "
import json
d=json.load(open('$O/bench_attn$v.json')); print('$v', round(d['value'],2), 'pairs/s', d['breakdown_ms'], d['clocks']['sm_mhz'], 'MHz', d['sample_scores'])"; done
echo "== ncu split stage"
VQA_ATTN_VARIANT=40 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc_d64 -s 2 -c 1 -o $O/attn_split python tools/bench_kernels.py attn-one > $O/ncu_attn.log 2>&1; echo "ncu rc=$?"
echo "== GEMM DRAM traffic, every GEMM launch of one step"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 -s 432 -c 432 --csv \
   --log-file $O/gemm_traffic.csv python bench.py --ncu > $O/gemm_traffic.log 2>&1; echo "ncu traffic rc=$?"
python tools/traffic_summary.py $O/gemm_traffic.csv $O/gemm_traffic.json > /dev/null 2>> $O/gemm_traffic.log
python -c "
import json
d=json.load(open('$O/gemm_traffic.json')); print('GEMM launches', d['launches'], 'read GB', round(d['dram_read_bytes']/1e9,1), 'write GB', round(d['dram_write_bytes']/1e9,1), 'ms', round(d['ms_under_ncu'],1)); print(d['by_template'])"
echo "== launch list of one step (time per kernel)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --ncu > $O/launches.log 2>&1; echo "ncu launches rc=$?"
echo "== default bench (CPU + HF baselines)"
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; cut -c1-3000 $O/bench_default.json
echo "== reference arm"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"; cut -c1-2500 $O/bench_reference.json
echo done
