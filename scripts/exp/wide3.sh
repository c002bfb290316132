#!/bin/bash
o=gpurun_out
run() {
  env "$@" python bench.py --workload $WL --no-e2e --no-cpu-baseline --others none > $o/w3.json 2> $o/w3.err
  python - <<PY
import json
d=json.load(open("$o/w3.json"))
print("$WL $*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
WL=dense_ovo_lambda
run ILLICO_WIDE_STAGES=6
run ILLICO_WIDE_STAGES=5
