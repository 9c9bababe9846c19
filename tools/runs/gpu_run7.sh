#!/bin/bash
# Qwen2.5-VL full-size bench + tests, DRAM-traffic sweep of the GEMM tile order, CLIP-T5 bench re-check.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run7.log
rm -f $L gpurun_out/check.jsonl gpurun_out/traffic_*.csv
run() { timeout 300 "$@" >> $L 2>&1; echo "rc=$? :: $*" >> $L; }
C="python tools/gpu_check.py"
run $C attention
run $C attention_perf 64 672 64 1
echo "== traffic sweep" >> $L
for G in 1024 2048 4096 8192; do
  for SH in "43008 20480 4096 gated_gelu" "43008 4096 10240 store" "43008 12288 4096 store" "43008 4096 4096 store"; do
    set -- $SH
    VQA_GEMM_GROUP_ROWS=$G timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
       -k regex:gemm_bf16 -s 2 -c 1 --csv --log-file gpurun_out/traffic_${G}_$1_$2_$3.csv $C gemm_once 2562 $1 $2 $3 $4 >> $L 2>&1
  done
done
for f in gpurun_out/traffic_*.csv; do echo "$f $(grep -E 'dram__bytes_read|dram__bytes_write|gpu__time|hit_rate' $f | awk -F'","' '{print $(NF-2)"="$NF}' | tr -d '"' | tr '\n' ' ')" >> $L; done
echo "== gemm perf by group rows (sustained loops)" >> $L
for G in 1024 2048 4096 8192; do
  VQA_GEMM_GROUP_ROWS=$G run $C gemm_perf 2562 43008 20480 4096 gated_gelu
  VQA_GEMM_GROUP_ROWS=$G run $C gemm_perf 2562 43008 4096 10240
done
echo "== pytest gpu" >> $L
timeout 1200 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
tail -3 gpurun_out/pytest_gpu.log >> $L
echo "== bench qwen" >> $L
timeout 900 python bench.py --model qwen2.5-vl-7b --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_qwen_r1.json 2>> $L; echo "bench qwen rc=$?" >> $L
cat gpurun_out/bench_qwen_r1.json >> $L
echo "== bench clipt5" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1f.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1f.json >> $L
grep -vE "^==PROF|^==WARN|^$" $L | cut -c1-600 | tail -70
