#!/bin/bash
# wide fused pass: parity tests, then the workloads it is for
o=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or golden or random or hand_back" > $o/wide_test.log 2>&1
tail -5 $o/wide_test.log
for wl in dense_ovo dense_ovo_lambda dense_ovo_highcount dense_ovo_continuous; do
  python bench.py --workload $wl --no-e2e --no-cpu-baseline --others none > $o/wide_$wl.json 2> $o/wide_$wl.err
  python - <<PY
import json
d=json.load(open("$o/wide_$wl.json"))
print("$wl", d["ms_per_step"], d["roofline"]["kernels_ms"])
PY
done
