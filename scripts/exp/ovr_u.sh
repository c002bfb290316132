#!/bin/bash
o=gpurun_out
python -m pytest tests -x -q -m gpu > $o/t_all.log 2>&1; tail -1 $o/t_all.log
ILLICO_OVR_HASH_MAX=16 ILLICO_OVR_TABLE=0 ILLICO_OVR_FUSED=0 python -m pytest tests -x -q -m gpu -k "golden or random or dispatchers_exact or float64 or k562_shape or tie_sum" > $o/ovr_test2.log 2>&1; tail -1 $o/ovr_test2.log
run() {
  env "$@" python bench.py --workload $WL --no-e2e --no-cpu-baseline --others none > $o/orc.json 2> $o/orc.err
  python - <<PY
import json
d=json.load(open("$o/orc.json"))
print("$WL $*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
WL=dense_ovr_unique
run A=1
WL=dense_ovo_unique
run A=1
WL=dense_ovo_continuous
run A=1
