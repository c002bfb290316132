#!/bin/bash
out=gpurun_out/exp_ab2.log
: > $out
cp illico_b200/libillico_b200.so /tmp/orig.so
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for v in v0 v1 v0 v1; do
  cp scripts/exp/variants/$v.so illico_b200/libillico_b200.so
  for wl in $WLS; do WL=$wl run V=$v WLN=$wl; done
done
cp /tmp/orig.so illico_b200/libillico_b200.so

cat $out
