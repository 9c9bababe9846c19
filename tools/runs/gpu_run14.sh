#!/bin/bash
# Round-1 run 14: L2 cache-hint policies on the GEMM operand streams: DRAM traffic of every GEMM launch of one step (ncu) for the auto
# policy vs no hints, then timing.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/run14.log
rm -f $L
timeout 600 python tools/gpu_check.py gemm 2562 512 4096 4096 store 0 1 >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -4 >> $L
for pol in auto 00; do
  if [ $pol = auto ]; then unset VQA_GEMM_L2_POLICY; else export VQA_GEMM_L2_POLICY=$pol; fi
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 -s 432 -c 432 --csv \
     --log-file gpurun_out/gemm_traffic_$pol.csv python bench.py --ncu >> $L 2>&1; echo "ncu traffic $pol rc=$?" >> $L
  python tools/traffic_summary.py gpurun_out/gemm_traffic_$pol.csv gpurun_out/gemm_traffic_$pol.json > /dev/null 2>> $L
done
for pol in auto 00 auto 00; do
  if [ $pol = auto ]; then unset VQA_GEMM_L2_POLICY; else export VQA_GEMM_L2_POLICY=$pol; fi
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pol_$pol.json 2>> $L
  python - <<PY >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_pol_$pol.json"))
print("bench policy=$pol", round(d["value"], 2), round(d["ms_per_step"], 2), d["breakdown_ms"], d["clocks"]["sm_mhz"], round(d["roofline"]["achieved"], 1))
PY
done
unset VQA_GEMM_L2_POLICY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:image_preprocess -s 3 -c 3 --csv \
   --log-file gpurun_out/preprocess_time.csv python tools/gpu_check.py preprocess_perf 64 512 336 >> $L 2>&1
grep image_preprocess gpurun_out/preprocess_time.csv | cut -d, -f5,10- | head -9 >> $L
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_preprocess -s 3 -c 1 -o gpurun_out/preprocess_full python tools/gpu_check.py preprocess_perf 64 512 336 >> $L 2>&1
python - <<'PY' >> $L 2>&1
import json
for pol in ("auto", "00"):
    d = json.load(open(f"gpurun_out/gemm_traffic_{pol}.json"))
    print(pol, "read GB", round(d["dram_read_bytes"] / 1e9, 1), "write GB", round(d["dram_write_bytes"] / 1e9, 1), "ms", round(d["ms_under_ncu"], 1))
    print("   ", d["largest_launches_sample"])
PY
grep -vE "^==PROF|^==WARN|^$|Warning|warn" $L | cut -c1-1200 | tail -30
