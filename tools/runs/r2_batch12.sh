#!/bin/bash
# round 2, GPU batch 12: split-K decoder GEMMs + query-tile split of the attention for small batches
set +e
O=gpurun_out/r2b12
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/kernels.log 2>&1; echo "rc=$?" >> $O/kernels.log; tail -3 $O/kernels.log
timeout 900 python -m pytest tests/test_gpu_clipt5.py -x -q > $O/clipt5.log 2>&1; echo "rc=$?" >> $O/clipt5.log; tail -3 $O/clipt5.log
for v in 30 40; do VQA_ATTN_VARIANT=$v timeout 300 python tools/bench_kernels.py attn >> $O/attn_ab.jsonl 2>> $O/attn_ab.err; done; cat $O/attn_ab.jsonl
for sk in 1 0 1 0; do
  VQA_GEMM_SPLITK=$sk timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_sk${sk}_$RANDOM.json 2>> $O/bench.err
done
for sk in 1 0; do for B in 1 4 16; do
  VQA_GEMM_SPLITK=$sk timeout 600 python bench.py --batch $B --graph --no-cpu-baseline --no-hf-baseline --steps 20 > $O/bench_b${B}_sk${sk}.json 2>> $O/bench.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b12/bench_*.json')):
    d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms', d['breakdown_ms'], d['clocks']['sm_mhz'], d['sample_scores'][:2])
PY
timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -s -k clipt5 > $O/fullwidth.log 2>&1; echo "rc=$?" >> $O/fullwidth.log; grep -E "spread|passed|failed|rc=" $O/fullwidth.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -3 $O/gpu_suite.log
echo done
