#!/bin/bash
o=gpurun_out/prof
mkdir -p $o
out=gpurun_out/exp_final4.log
: > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $out
for wl in dense_ovo dense_ovr csr_ovo csr_ovr; do
  timeout 600 python bench.py --workload $wl --continuous --no-e2e --no-cpu-baseline --steps 3 > $o/bench_continuous_$wl.json 2>/dev/null
done
ILLICO_OVO_FUSED=0 timeout 600 python bench.py --workload dense_ovo --no-e2e --no-cpu-baseline > $o/bench_general_dense_ovo.json 2>/dev/null
ILLICO_OVO_FUSED=0 timeout 600 python bench.py --workload csr_ovo --no-e2e --no-cpu-baseline > $o/bench_general_csr_ovo.json 2>/dev/null
for f in continuous_dense_ovo continuous_dense_ovr continuous_csr_ovo continuous_csr_ovr general_dense_ovo general_csr_ovo; do python - $o/bench_$f.json $f >> $out <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print(sys.argv[2], 'ms', d['ms_per_step'], 'stage', r['stage_ms'], 'rank', r['rank_ms'], 'fused', r.get('fused_ms'))
PY
done
cat $out
