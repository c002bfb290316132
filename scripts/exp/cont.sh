#!/bin/bash
out=gpurun_out/exp_cont.log
: > $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 600 python bench.py --workload $WL --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'], 'fused', r.get('fused_ms'), 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
WL="csr_ovo --continuous" run A=1
WL="csr_ovr --continuous" run A=1
WL=csr_ovo run A=1
timeout 900 python -m pytest tests -m gpu -x -q -k "csr or random" 2>&1 | tail -3 >> $out
cat $out
