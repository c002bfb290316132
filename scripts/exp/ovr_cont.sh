#!/bin/bash
o=gpurun_out
python -m pytest tests -x -q -m gpu -k "golden or random or dispatchers_exact or float64 or k562_shape or tie_sum" > $o/ovr_test.log 2>&1
tail -2 $o/ovr_test.log
ILLICO_OVR_HASH_MAX=64 python -m pytest tests -x -q -m gpu -k "golden or random or dispatchers_exact or float64 or k562_shape or tie_sum" > $o/ovr_test2.log 2>&1
tail -2 $o/ovr_test2.log
run() {
  env "$@" python bench.py --workload $WL --no-e2e --no-cpu-baseline --others none > $o/orc.json 2> $o/orc.err
  python - <<PY
import json
d=json.load(open("$o/orc.json"))
print("$WL $*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
WL=dense_ovr_continuous
run ILLICO_OVR_HASH_MAX=2048
run ILLICO_OVR_HASH_MAX=256
run ILLICO_OVR_HASH_MAX=256 ILLICO_OVR_BUCKETS=0
