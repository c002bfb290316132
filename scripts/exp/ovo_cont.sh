#!/bin/bash
# general one-versus-reference path on continuous data: parity, then step times with / without the bucket index
o=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or random or dispatchers_exact or stream_tier or fused_ovo or two_million" > $o/ovo_test.log 2>&1
tail -3 $o/ovo_test.log
run() {
  env "$@" python bench.py --workload $WL --no-e2e --no-cpu-baseline --others none > $o/oc.json 2> $o/oc.err
  python - <<PY
import json
d=json.load(open("$o/oc.json"))
print("$WL $*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
WL=dense_ovo_continuous
run A=1
run ILLICO_OVO_BUCKETS=0
WL=csr_ovo_continuous
run A=1
run ILLICO_OVO_BUCKETS=0
