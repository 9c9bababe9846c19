#!/bin/bash
set +e
O=gpurun_out/r2ab
mkdir -p $O
for i in 1 2 3; do
  (cd .r1_ab && timeout 600 python bench.py --no-cpu-baseline > ../$O/r1_$i.json 2>> ../$O/err.log)
  timeout 600 python bench.py --no-cpu-baseline --no-hf-baseline > $O/r2_$i.json 2>> $O/err.log
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2ab/r*.json')):
    d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],2), 'pairs/s', round(d['ms_per_step'],2), d['breakdown_ms'], d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
PY
tail -3 $O/err.log
