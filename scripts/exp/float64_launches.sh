#!/bin/bash
o=gpurun_out
python scripts/exp/float64_run.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/launches_float64.csv python scripts/exp/float64_run.py > $o/launches_float64.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_float64.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
ci={h:i for i,h in enumerate(rows[hdr])}
acc=collections.Counter()
for r in rows[hdr+1:]:
    acc[r[ci['Kernel Name']].split('(')[0][:110]]+=1
for k,v in sorted(acc.items(), key=lambda kv:-kv[1]): print(f"{v:5d}  {k}")
PY
