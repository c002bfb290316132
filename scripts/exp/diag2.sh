#!/bin/bash
out=gpurun_out/exp_diag2.log
: > $out
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'], 'frac', r['frac'])
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for al in 8 32; do
for d in 0 2 3 4 5; do
  run ILLICO_B200_SLOT_ALIGN=$al ILLICO_STAGE_TMA_DIAG=$d ILLICO_STAGE_TMA_CFG=5
done
done
run ILLICO_B200_SLOT_ALIGN=32 ILLICO_STAGE_TMA_DIAG=3 ILLICO_STAGE_TMA_CFG=0
run ILLICO_B200_SLOT_ALIGN=32 ILLICO_STAGE_TMA_DIAG=2 ILLICO_STAGE_TMA_CFG=0
cat $out
