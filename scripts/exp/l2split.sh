#!/bin/bash
out=gpurun_out/exp_l2split.log
: > $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 300 python bench.py --workload $WL --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'frac', r['frac'])
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for sp in 0 1 2 3 4; do WL=csr_ovo run ILLICO_CSR_L2_SPLIT=$sp; done
ILLICO_CSR_L2_SPLIT=2 timeout 600 python -m pytest tests -m gpu -x -q -k "csr or random or golden" 2>&1 | tail -3 >> $out
cat $out
