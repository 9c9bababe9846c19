#!/bin/bash
# round 2, GPU batch 8: shipped build (split-row attention default): smoke, full GPU suite, default bench with baselines, config-4 per-rank share,
# launch list + GEMM traffic of one step for profiles/
set +e
O=gpurun_out/r2b8
mkdir -p $O
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log | cut -c1-400
echo "== full GPU suite"; timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -4 $O/gpu_suite.log
echo "== default bench (CPU + HF baselines)"
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; cut -c1-1200 $O/bench_default.json
echo "== config 4 per-rank share: job of 1250 pairs on one GPU"
timeout 900 python bench.py --pairs 1250 --steps 2 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_pairs1250.json 2> $O/bench_pairs1250.err; echo "rc=$?"; cut -c1-1500 $O/bench_pairs1250.json
echo "== ragged"
timeout 900 python bench.py --ragged --no-cpu-baseline --no-hf-baseline > $O/bench_ragged.json 2> $O/bench_ragged.err; echo "rc=$?"; cut -c1-600 $O/bench_ragged.json
echo "== launch list of one step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --ncu > $O/launches.log 2>&1; echo "ncu launches rc=$?"
echo "== GEMM DRAM traffic, every GEMM launch of one step"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 -s 432 -c 432 --csv \
   --log-file $O/gemm_traffic.csv python bench.py --ncu > $O/gemm_traffic.log 2>&1; echo "ncu traffic rc=$?"
python tools/traffic_summary.py $O/gemm_traffic.csv $O/gemm_traffic.json > /dev/null 2>> $O/gemm_traffic.log
echo "== ncu --set full of the hot GEMM (encoder wi) and the attention kernel inside the step"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_sm100_kernel -s 529 -c 1 -o $O/gemm_wi python bench.py --ncu > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
echo done
