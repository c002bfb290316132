#!/bin/bash
o=gpurun_out
run() {
  env "$@" python bench.py --workload $WL --no-e2e --no-cpu-baseline --others none > $o/un.json 2> $o/un.err
  python - <<PY
import json
d=json.load(open("$o/un.json"))
print("$WL $*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
WL=dense_ovr_unique
run A=1
run ILLICO_OVR_BUCKETS=0
WL=dense_ovo_unique
run A=1
run ILLICO_OVO_BUCKETS=0
