#!/bin/bash
out=gpurun_out/exp_fused2.log
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
for c in 0 1 2 3 4; do run ILLICO_OVO_FUSED_CFG=$c; done
run ILLICO_OVO_FUSED_CFG=0 ILLICO_OVO_FUSED_ROWS=768
run ILLICO_OVO_FUSED_CFG=3 ILLICO_OVO_FUSED_ROWS=768
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
cat $out
