#!/bin/bash
# in-call A/B of library variants (scripts/exp/variants/*.so): boxes differ by a few percent, variants must share one
out=gpurun_out/exp_ab.log
: > $out
cp illico_b200/libillico_b200.so /tmp/orig.so
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'])
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for rep in 1 2; do
for v in $VARIANTS; do
  cp scripts/exp/variants/$v.so illico_b200/libillico_b200.so
  for c in $CFGS; do run V=$v ILLICO_OVO_FUSED_CFG=$c; done
done
done
cp /tmp/orig.so illico_b200/libillico_b200.so
cat $out
