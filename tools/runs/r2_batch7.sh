#!/bin/bash
# round 2, GPU batch 7: does the mbarrier suspend-time hint in the GEMM's waits change its DRAM traffic / time? A/B of two builds on one box
set +e
O=gpurun_out/r2b7
mkdir -p $O
SPIN=$PWD/t2v_metrics_b200/libvqa_b200_gemmspin.so
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_bf16 --csv"
for tag in hint spin hint spin; do
  if [ $tag = spin ]; then export VQA_B200_LIB=$SPIN; else unset VQA_B200_LIB; fi
  SCHEDS=0:0 timeout 600 ncu $M --log-file $O/iso_${tag}_$RANDOM.csv python tools/bench_kernels.py gemm-ncu > /dev/null 2>&1
  SCHEDS=0:0 timeout 600 python tools/bench_kernels.py gemm-time >> $O/time_$tag.jsonl 2>> $O/time.err
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hf-baseline > $O/bench_${tag}_$RANDOM.json 2>> $O/bench.err
done
unset VQA_B200_LIB
python - <<'PY'
import csv, glob, json
for f in sorted(glob.glob('gpurun_out/r2b7/iso_*.csv')):
    lines=[l for l in open(f) if l.startswith('"')]
    per={}
    for r in csv.DictReader(lines):
        d=per.setdefault(int(r['ID']), {})
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']; n=r['Metric Name']
        mul=dict(byte=1,Kbyte=1e3,Mbyte=1e6,Gbyte=1e9,ns=1e-6,us=1e-3,ms=1,s=1e3).get(u,1)
        d[n.split('.')[0][-18:]]=v*mul
    print(f, ' | '.join(f"rd {d['dram__bytes_read']/1e9:.2f}G {d['gpu__time_duration']:.3f}ms hit {d['_t_sector_hit_rate']:.0f}%" for k,d in sorted(per.items())))
for f in sorted(glob.glob('gpurun_out/r2b7/time_*.jsonl')):
    for l in open(f):
        print(f, [(r['shape'], r['ms'], r['tflops']) for r in json.loads(l)])
for f in sorted(glob.glob('gpurun_out/r2b7/bench_*.json')):
    d=json.load(open(f)); print(f, round(d['value'],2), d['breakdown_ms'], d['clocks']['sm_mhz'])
PY
echo done
