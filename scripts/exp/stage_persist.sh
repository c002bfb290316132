#!/bin/bash
# dense staging: 2-D grid against the persistent one, register budgets (general path forced, continuous data)
o=gpurun_out
run() {
  env "$@" ILLICO_OVO_FUSED=0 python bench.py --workload dense_ovo_continuous --no-e2e --no-cpu-baseline --others none > $o/sp.json 2> $o/sp.err
  python - <<PY
import json
d=json.load(open("$o/sp.json"))
print("$*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
for cfg in 0 4 5; do
run ILLICO_STAGE_TMA_CFG=$cfg
run ILLICO_STAGE_TMA_CFG=$cfg ILLICO_STAGE_PERSIST=1
done
