#!/bin/bash
out=gpurun_out/exp_fused.log
: > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $out
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --workload ${WL:-dense_ovo} --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'], 'frac', r['frac'], 'launches', d.get('gpu_launches'))
    elif 'Warning' not in l and 'to_sparse' not in l: print(l.rstrip())
" >> $out
}
run ILLICO_OVO_FUSED=1 ILLICO_OVO_FUSED_CFG=0
run ILLICO_OVO_FUSED=1 ILLICO_OVO_FUSED_CFG=1
run ILLICO_OVO_FUSED=1 ILLICO_OVO_FUSED_CFG=2
run ILLICO_OVO_FUSED=1 ILLICO_OVO_FUSED_CFG=0 ILLICO_OVO_FUSED_ROWS=768
run ILLICO_OVO_FUSED=1 ILLICO_OVO_FUSED_CFG=0 ILLICO_OVO_FUSED_ROWS=3072
ILLICO_OVO_FUSED_CFG=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ovo_fused_kernel -c 1 -o gpurun_out/prof_fused python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/prof_fused.log 2>&1
cat $out
