#!/bin/bash
# round 2, GPU batch 9: library after the clean-up (two-pass stage removed): GPU suite; Qwen launch list (where does the Qwen step go?)
set +e
O=gpurun_out/r2b9
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/gpu_suite.log 2>&1; echo "rc=$?" >> $O/gpu_suite.log; tail -4 $O/gpu_suite.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/qwen_launches.csv python bench.py --model qwen2.5-vl-7b --ncu > $O/qwen_launches.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/r2b9/qwen_launches.csv') if l.startswith('"')]
rows=[r for r in csv.DictReader(lines) if r['Metric Name']=='gpu__time_duration.sum' and 'vqa::' in r['Kernel Name']]
n=len(rows)//2
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[n:]:
    v=float(r['Metric Value'].replace(',',''))*dict(ns=1e-6,us=1e-3,ms=1,s=1e3).get(r['Metric Unit'],1e-6)
    k=r['Kernel Name'].split('(')[0].replace('void ','').replace('vqa::','')+' grid='+r['Grid Size']
    agg[k][0]+=1; agg[k][1]+=v
print('launches per step', n, 'sum ms', round(sum(v for _,v in agg.values()),2))
for k,(c,ms) in sorted(agg.items(), key=lambda x:-x[1][1])[:28]:
    print(f"{k[:100]:100s} {c:4d} {ms:8.3f}")
PY
timeout 600 python bench.py --batch 1 --no-cpu-baseline --no-hf-baseline --steps 20 > $O/bench_b1.json 2> $O/bench_b1.err; timeout 600 python bench.py --batch 1 --graph --no-cpu-baseline --no-hf-baseline --steps 20 > $O/bench_b1_graph.json 2> $O/bench_b1_graph.err
python -c "
import json
for f in ('bench_b1','bench_b1_graph'):
    d=json.load(open('gpurun_out/r2b9/'+f+'.json')); print(f, round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms/step')"
echo done
