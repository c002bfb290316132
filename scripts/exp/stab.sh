#!/bin/bash
out=gpurun_out/exp_stab.log
: > $out
ILLICO_OVO_SEARCH_TABLE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 300 python bench.py --workload $WL --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'rank_ms', r['rank_ms'], 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
WL="dense_ovo --high-count-frac 0.02" run ILLICO_OVO_SEARCH_TABLE=1
WL="dense_ovo --high-count-frac 0.2" run ILLICO_OVO_SEARCH_TABLE=1
WL="dense_ovo --continuous" run ILLICO_OVO_SEARCH_TABLE=1
WL="dense_ovo --continuous" run ILLICO_OVO_SEARCH_TABLE=0
cat $out
