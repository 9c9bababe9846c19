#!/bin/bash
# round 2, GPU batch 11: where does a batch-1 CLIP-FlanT5 call go? (launch list at B = 1 and B = 4)
set +e
O=gpurun_out/r2b11
mkdir -p $O
for B in 1 4; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_b$B.csv python bench.py --ncu --batch $B > $O/launches_b$B.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections
lines=[l for l in open('$O/launches_b$B.csv') if l.startswith('"')]
rows=[r for r in csv.DictReader(lines) if r['Metric Name']=='gpu__time_duration.sum' and 'vqa::' in r['Kernel Name']]
n=len(rows)//2
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[n:]:
    v=float(r['Metric Value'].replace(',',''))*dict(ns=1e-6,us=1e-3,ms=1,s=1e3).get(r['Metric Unit'],1e-6)
    k=r['Kernel Name'].split('(')[0].replace('void ','').replace('vqa::','')+' grid='+r['Grid Size']
    agg[k][0]+=1; agg[k][1]+=v
print('B=$B launches per step', n, 'sum ms', round(sum(v for _,v in agg.values()),3))
for k,(c,ms) in sorted(agg.items(), key=lambda x:-x[1][1])[:22]:
    print(f"{k[:90]:90s} {c:4d} {ms:8.3f} {1000*ms/c:8.1f} us")
PY
done
for B in 1 4 16; do
timeout 600 python bench.py --batch $B --graph --no-cpu-baseline --no-hf-baseline --steps 20 > $O/bench_b${B}_graph.json 2> $O/bench_b${B}_graph.err
python -c "
import json
d=json.load(open('$O/bench_b${B}_graph.json')); print('B=$B graph', round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms/step', d['breakdown_ms'])"
done
echo done
