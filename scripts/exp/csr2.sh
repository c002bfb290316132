#!/bin/bash
out=gpurun_out/exp_csr2.log
: > $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-csr_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'kernel', r['kernel'], 'frac', r['frac'])
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
WL=csr_ovo run A=1
WL=csr_ovr run A=1
timeout 600 python -m pytest tests -m gpu -x -q -k "csr or random or golden" 2>&1 | tail -3 >> $out
cat $out
