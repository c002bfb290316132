#!/bin/bash
# TMA dense staging: parity + ring configurations (results: gpurun_out/exp_stage_tma.log)
out=gpurun_out/exp_stage_tma.log
: > $out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $out
for cfg in off 0 1 2 3 4 5; do
  echo "== ILLICO_STAGE_TMA_CFG=$cfg" >> $out
  if [ $cfg = off ]; then export ILLICO_STAGE_TMA=0; else export ILLICO_STAGE_TMA=1 ILLICO_STAGE_TMA_CFG=$cfg; fi
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'], 'frac', r['frac'])
    else: print(l.rstrip())
" >> $out
done
cat $out
