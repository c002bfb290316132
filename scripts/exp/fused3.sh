#!/bin/bash
out=gpurun_out/exp_fused3.log
: > $out
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'frac', r['frac'], 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for c in 0 3 0 3; do run ILLICO_OVO_FUSED_CFG=$c; done
cat $out
