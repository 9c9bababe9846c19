#!/bin/bash
# Second pass: full GPU test suite, smoke, first full-size bench line, ncu launch list + one full capture of the top kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run2.log
rm -f $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi2.txt 2>&1
nproc >> gpurun_out/nvsmi2.txt; free -g | head -2 >> gpurun_out/nvsmi2.txt
echo "== pytest gpu" >> $L
timeout 1200 python -m pytest tests/ -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $L
tail -40 gpurun_out/pytest_gpu.log >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "smoke rc=$?" >> $L
echo "== bench" >> $L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2>> $L; echo "bench rc=$?" >> $L
cat gpurun_out/bench_r1.json >> $L
echo "== ncu launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 629 -c 629 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --ncu >> $L 2>&1; echo "ncu list rc=$?" >> $L
echo "== ncu full (top GEMM kernel)" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 -s 900 -c 3 -o gpurun_out/prof_gemm_r1 \
    python bench.py --ncu >> $L 2>&1; echo "ncu full rc=$?" >> $L
tail -30 $L
