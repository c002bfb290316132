#!/bin/bash
out=gpurun_out/exp_ovo.log
: > $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $out
for wl in dense_ovo csr_ovo; do
  echo "== $wl" >> $out
  python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'])
    else: print(l.rstrip())
" >> $out
done
cat $out
