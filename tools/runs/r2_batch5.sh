#!/bin/bash
# round 2, GPU batch 5: why do the GEMMs read ~2x more DRAM inside the step than the same launch in isolation? (ncu dram bytes, several set-ups)
set +e
O=gpurun_out/r2b5
mkdir -p $O
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_bf16 --csv"
echo "== E0 isolated launches, automatic order"
SCHEDS=0:0 timeout 600 ncu $M --log-file $O/iso.csv python tools/bench_kernels.py gemm-ncu > $O/iso.log 2>&1; echo rc=$?
echo "== E3 isolated + 60 GB ballast"
SCHEDS=0:0 BALLAST_GB=60 timeout 600 ncu $M --log-file $O/iso_ballast.csv python tools/bench_kernels.py gemm-ncu > $O/iso_ballast.log 2>&1; echo rc=$?
echo "== E4 isolated, zero operands"
SCHEDS=0:0 ZERO_OPERANDS=1 timeout 600 ncu $M --log-file $O/iso_zero.csv python tools/bench_kernels.py gemm-ncu > $O/iso_zero.log 2>&1; echo rc=$?
echo "== E1 in-step, layer 0..2 of the encoder (launches 92..108 of step 2), cache-control all vs none"
timeout 900 ncu $M -s 524 -c 20 --log-file $O/step_all.csv python bench.py --ncu > $O/step_all.log 2>&1; echo rc=$?
timeout 900 ncu $M --cache-control none -s 524 -c 20 --log-file $O/step_none.csv python bench.py --ncu > $O/step_none.log 2>&1; echo rc=$?
echo "== E2 in-step, round-1 order"
VQA_GEMM_SCHEDULE=r1 timeout 900 ncu $M -s 524 -c 20 --log-file $O/step_r1.csv python bench.py --ncu > $O/step_r1.log 2>&1; echo rc=$?
echo "== E5 in-step at batch 16 (M = 10752)"
timeout 900 ncu $M -s 524 -c 20 --log-file $O/step_b16.csv python bench.py --ncu --batch 16 > $O/step_b16.log 2>&1; echo rc=$?
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/r2b5/*.csv')):
    lines=[l for l in open(f) if l.startswith('"')]
    per={}
    for r in csv.DictReader(lines):
        d=per.setdefault(int(r['ID']), dict(name=r['Kernel Name'].split('(')[0][-12:]))
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']; n=r['Metric Name']
        mul=dict(byte=1,Kbyte=1e3,Mbyte=1e6,Gbyte=1e9,ns=1e-6,us=1e-3,ms=1,s=1e3).get(u,1)
        d[n.split('.')[0][-24:]]=v*mul
    print(f)
    for k in sorted(per):
        d=per[k]; print('  ',k,d['name'],' '.join(f"{a}={b/1e9:.2f}G" if 'bytes' in a else (f"{a}={b:.3f}" ) for a,b in d.items() if a!='name'))
PY
echo done
