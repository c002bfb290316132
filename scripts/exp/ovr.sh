#!/bin/bash
out=gpurun_out/exp_ovrf.log
: > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 >> $out
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovr} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'], 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
run ILLICO_OVR_FUSED=1
run ILLICO_OVR_FUSED=0
WL=dense_ovo run ILLICO_OVO_FUSED=1
cat $out
