#!/bin/bash
# one-off tuning sweep of the dense staging kernel's cache policies (results: gpurun_out/exp_stage_policy.log)
out=gpurun_out/exp_stage_policy.log
: > $out
for pol in 0 1 2 3 4 5 6 7 8; do
  echo "== ILLICO_STAGE_POLICY=$pol" >> $out
  ILLICO_STAGE_POLICY=$pol python bench.py --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'])
    else: print(l.rstrip())
" >> $out
done
cat $out
