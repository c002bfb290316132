#!/bin/bash
out=gpurun_out/exp_diag.log
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
run ILLICO_STAGE_TMA_DIAG=1 ILLICO_STAGE_TMA_CFG=0
run ILLICO_STAGE_TMA_DIAG=1 ILLICO_STAGE_TMA_CFG=5
run ILLICO_STAGE_TMA_DIAG=1 ILLICO_STAGE_TMA_CFG=1
for al in 8 16 32; do
  WL=dense_ovo run ILLICO_B200_SLOT_ALIGN=$al
  WL=dense_ovr run ILLICO_B200_SLOT_ALIGN=$al
done
WL=csr_ovo run ILLICO_B200_SLOT_ALIGN=32
ILLICO_B200_SLOT_ALIGN=32 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $out
cat $out
