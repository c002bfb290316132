#!/bin/bash
out=gpurun_out/exp_csr3.log
: > $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-csr_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'frac', r['frac'])
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for m in 0 1 2 3; do WL=csr_ovo run ILLICO_CSR_COUNTS_KERNEL=$m; done
WL=csr_ovr run ILLICO_CSR_COUNTS_KERNEL=3
timeout 900 python -m pytest tests -m gpu -x -q -k "csr or random or golden or dispatch" 2>&1 | tail -4 >> $out
cat $out
