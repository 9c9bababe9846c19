#!/bin/bash
# round 2, GPU batch 2: rounding-point fixes (parity), attention diagnostics (phases, ncu), GEMM tile-order A/B in the real step, full GPU suite
set +e
O=gpurun_out/r2b2
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi.txt 2>&1
echo "== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -s > $O/kernels.log 2>&1; echo "rc=$?" >> $O/kernels.log; grep -E "round_scores kernel|passed|failed|rc=" $O/kernels.log | tail -8
echo "== full-width parity"; timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s > $O/fullwidth.log 2>&1; echo "rc=$?" >> $O/fullwidth.log; grep -E "^\[|passed|failed|Error|rc=" $O/fullwidth.log | tail -40
echo "== attention A/B (d64)"
for v in 1 20 21; do VQA_ATTN_VARIANT=$v timeout 300 python tools/bench_kernels.py attn >> $O/attn_ab.jsonl 2>> $O/attn_ab.err; done
cat $O/attn_ab.jsonl
echo "== attention phases"; timeout 300 python tools/bench_kernels.py attn-phases > $O/attn_phases.json 2> $O/attn_phases.err; cat $O/attn_phases.json
echo "== attention d128 A/B"
for v in 1 10 12; do VQA_ATTN128_VARIANT=$v timeout 300 python tools/bench_kernels.py attn128 >> $O/attn128_ab.jsonl 2>> $O/attn128_ab.err; done
cat $O/attn128_ab.jsonl
echo "== ncu attention (full set, source counters)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc_d64_kernel -s 2 -c 1 -o $O/attn_v2 python tools/bench_kernels.py attn-one > $O/ncu_attn.log 2>&1; echo "ncu rc=$?"
echo "== bench A/B: round-1 tile order vs this round's"
VQA_GEMM_SCHEDULE=r1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_sched_r1.json 2> $O/bench_sched_r1.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_sched_r2.json 2> $O/bench_sched_r2.err
VQA_GEMM_SCHEDULE=r1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_sched_r1b.json 2> $O/bench_sched_r1b.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_sched_r2b.json 2> $O/bench_sched_r2b.err
for f in r1 r2 r1b r2b; do python -c "
import json,sys
d=json.load(open('$O/bench_sched_$f.json')); print('$f', round(d['value'],2), 'pairs/s', d['breakdown_ms'], d['clocks']['sm_mhz'], 'MHz')"; done
echo "== qwen bench"; timeout 900 python bench.py --model qwen2.5-vl-7b --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_qwen.json 2> $O/bench_qwen.err; head -c 2500 $O/bench_qwen.json; echo
echo "== full GPU suite"; timeout 2400 python -m pytest tests/ -q -m gpu --deselect tests/test_gpu_fullwidth.py > $O/suite.log 2>&1; echo "rc=$?" >> $O/suite.log; tail -15 $O/suite.log
echo done
