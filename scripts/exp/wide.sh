#!/bin/bash
# wide fused pass: parity tests, then the workloads it is for
o=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or golden or random or hand_back" > $o/wide_test.log 2>&1
tail -3 $o/wide_test.log
run() {
  python bench.py --workload $1 --no-e2e --no-cpu-baseline --others none > $o/wide_$1$2.json 2> $o/wide_$1$2.err
  python - <<PY
import json
d=json.load(open("$o/wide_$1$2.json"))
print("$1$2", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
for wl in dense_ovo dense_ovo_lambda dense_ovo_highcount; do run $wl; done
ILLICO_FUSED_BACKOFF=1 run dense_ovo _backoff
ILLICO_FUSED_BACKOFF=1 run dense_ovr _backoff
run dense_ovr
run dense_ovo _again
