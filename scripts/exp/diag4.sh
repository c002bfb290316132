#!/bin/bash
out=gpurun_out/exp_diag4.log
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
for pad in 0 32 96 224 992 4064 16352; do
  run ILLICO_B200_SLOT_ALIGN=32 ILLICO_B200_SLOT_PAD=$pad ILLICO_STAGE_TMA_CFG=5
done
for rows in 256 1024 2048; do
  run ILLICO_B200_SLOT_ALIGN=32 ILLICO_STAGE_ROWS=$rows ILLICO_STAGE_TMA_CFG=5
done
cat $out
