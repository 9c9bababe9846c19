#!/bin/bash
# round 2, GPU batch 3: mbarrier suspend hint + streaming softmax stage (A/B), parity diagnostics (cross-attention association, sdpa reference)
set +e
O=gpurun_out/r2b3
mkdir -p $O
echo "== kernel tests (default stage, then streaming stage)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > $O/kernels.log 2>&1; echo "rc=$?" >> $O/kernels.log; tail -2 $O/kernels.log
VQA_ATTN_VARIANT=31 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/kernels_stream_round.log 2>&1; echo "rc=$?" >> $O/kernels_stream_round.log; tail -2 $O/kernels_stream_round.log
VQA_ATTN_VARIANT=30 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention and not rounding" > $O/kernels_stream.log 2>&1; echo "rc=$?" >> $O/kernels_stream.log; tail -2 $O/kernels_stream.log
echo "== attention A/B (d64): 20/21 two-pass stage, 30/31 streaming stage (x0 fp32 scores, x1 reference rounding)"
for v in 20 21 30 31; do VQA_ATTN_VARIANT=$v ATTN_ROUND=$((v % 10)) timeout 300 python tools/bench_kernels.py attn >> $O/attn_ab.jsonl 2>> $O/attn_ab.err; done
cat $O/attn_ab.jsonl
echo "== attention phases (two-pass stage)"; timeout 300 python tools/bench_kernels.py attn-phases > $O/attn_phases.json 2> $O/attn_phases.err; cat $O/attn_phases.json
echo "== clipt5 goldens with the streaming stage"; VQA_ATTN_VARIANT=31 timeout 900 python -m pytest tests/test_gpu_clipt5.py -x -q > $O/clipt5_stream.log 2>&1; echo "rc=$?" >> $O/clipt5_stream.log; tail -3 $O/clipt5_stream.log
echo "== full-width parity"; timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s -k "clipt5" > $O/fullwidth.log 2>&1; echo "rc=$?" >> $O/fullwidth.log; grep -E "^\[|passed|failed|Error|rc=" $O/fullwidth.log | tail -40
echo "== ncu streaming stage"
VQA_ATTN_VARIANT=31 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc_d64 -s 2 -c 1 -o $O/attn_stream python tools/bench_kernels.py attn-one > $O/ncu_attn.log 2>&1; echo "ncu rc=$?"
echo "== bench: two-pass vs streaming stage in the full step"
VQA_ATTN_VARIANT=21 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_attn21.json 2> $O/bench_attn21.err
VQA_ATTN_VARIANT=31 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_attn31.json 2> $O/bench_attn31.err
for f in 21 31; do python -c "
import json
d=json.load(open('$O/bench_attn$f.json')); print('$f', round(d['value'],2), 'pairs/s', d['breakdown_ms'], d['clocks']['sm_mhz'], 'MHz', d['sample_scores'])"; done
echo "== fused norms: kernel hooks, engine modes, bench"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k fused_rmsnorm > $O/normfuse_kernels.log 2>&1; echo "rc=$?" >> $O/normfuse_kernels.log; tail -3 $O/normfuse_kernels.log
timeout 900 python -m pytest tests/test_gpu_clipt5.py -x -q -s -k "modes" > $O/modes.log 2>&1; echo "rc=$?" >> $O/modes.log; grep -E "^\[|passed|failed|rc=" $O/modes.log | tail -12
for cfgs in "0 0" "1 0" "0 1"; do set -- $cfgs; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline --fuse-norms $1 --round-scores $2 > $O/bench_f$1_r$2.json 2> $O/bench_f$1_r$2.err
python -c "
import json
d=json.load(open('$O/bench_f$1_r$2.json')); print('fuse=$1 round=$2', round(d['value'],2), 'pairs/s', d['breakdown_ms'], d['clocks']['sm_mhz'], 'MHz', d['sample_scores'])"; done
echo done
