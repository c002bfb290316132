#!/bin/bash
out=gpurun_out/exp_final2.log
: > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $out; timeout 600 python -m pytest scripts/exp/test_compact.py -x -q -p no:cacheprovider 2>&1 | tail -8 >> $out
run() {
  echo "== $WL $*" >> $out
  env "$@" timeout 300 python bench.py --workload $WL --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'fused_ms', r.get('fused_ms'), 'frac', r['frac'], 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
for sp in 0 2 0 2 1 3; do WL=csr_ovo run ILLICO_CSR_L2_SPLIT=$sp; done
WL="dense_ovo --high-count-frac 0.02" run ILLICO_FUSED_COMPACT=1
WL="dense_ovo --high-count-frac 0.02" run ILLICO_FUSED_COMPACT=0 ILLICO_FUSED_MAX_HANDBACK=1.0
WL="dense_ovo --high-count-frac 0.02" run ILLICO_OVO_FUSED=0
WL="dense_ovr --high-count-frac 0.05" run ILLICO_FUSED_COMPACT=1
WL="dense_ovr --high-count-frac 0.05" run ILLICO_OVR_FUSED=0
cat $out
