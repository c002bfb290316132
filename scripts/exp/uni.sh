#!/bin/bash
out=gpurun_out/exp_uni.log
: > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 >> $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovr} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'frac', r['frac'], 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for st in 4 5; do
WL=dense_ovo run ILLICO_FUSED_STAGES=$st
WL=dense_ovr run ILLICO_FUSED_STAGES=$st
done
WL=dense_ovo run ILLICO_FUSED_STAGES=4 ILLICO_FUSED_ROWS=768
WL=dense_ovo run ILLICO_FUSED_STAGES=4 ILLICO_FUSED_ROWS=3072
cat $out
