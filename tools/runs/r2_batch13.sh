#!/bin/bash
# round 2, GPU batch 13: programmatic dependent launch (GEMM / attention kernels) A/B; secondary video shape 8x24x24
set +e
O=gpurun_out/r2b13
mkdir -p $O
for i in 1 2; do timeout 1800 python -m pytest tests -m gpu -q -x > $O/gpu_suite_$i.log 2>&1; echo "rc=$?" >> $O/gpu_suite_$i.log; tail -2 $O/gpu_suite_$i.log | head -1; done
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
for pdl in 1 0 1 0; do
  VQA_PDL=$pdl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_pdl${pdl}_$RANDOM.json 2>> $O/bench.err
done
for pdl in 1 0; do for B in 1 4 16; do
  VQA_PDL=$pdl timeout 600 python bench.py --batch $B --graph --no-cpu-baseline --no-hf-baseline --steps 20 > $O/bench_b${B}_pdl${pdl}.json 2>> $O/bench.err
done; done
for pdl in 1 0; do VQA_PDL=$pdl timeout 900 python bench.py --model qwen2.5-vl-7b --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_pdl${pdl}.json 2>> $O/bench.err; done
timeout 900 python bench.py --model qwen2.5-vl-7b --video --video-size 336 --no-cpu-baseline --no-hf-baseline > $O/bench_qwen_video336.json 2>> $O/bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b13/bench_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms', d['breakdown_ms'], d['clocks']['sm_mhz'], d['sample_scores'][:2])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 $O/bench.err
echo done
