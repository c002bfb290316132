#!/bin/bash
out=gpurun_out/exp_final5.log
: > $out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 200 python bench.py --workload $WL --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'rank_ms', r['rank_ms'])
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
WL="dense_ovo --high-count-frac 0.02" run A=1
WL="dense_ovo --continuous" run A=1
cat $out
