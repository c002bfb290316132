#!/bin/bash
# state check: GPU tests + the four K562 workloads (kernel-only) in one call
out=gpurun_out/state.log
: > $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $out
for wl in dense_ovo dense_ovr csr_ovo csr_ovr; do
  echo "== $wl" >> $out
  python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('ms_per_step', d['ms_per_step'], 'stage_ms', r['stage_ms'], 'rank_ms', r['rank_ms'], 'frac', r['frac'])
    else: print(l.rstrip())
" >> $out
done
cat $out
