#!/bin/bash
# CSR fused pass without shared-memory atomics (warp = gene range) against the atomic one
o=gpurun_out
mkdir -p $o/prof3
run() {
  env "$@" python bench.py --workload $WL --no-e2e --no-cpu-baseline --others none > $o/c2.json 2> $o/c2.err
  python - <<PY
import json
d=json.load(open("$o/c2.json"))
print("$WL $*", d["ms_per_step"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if v > 0.03})
PY
}
for WL in csr_ovo; do
run ILLICO_CSR_PASS=2
run ILLICO_CSR_PASS=4
run ILLICO_CSR_PASS=1
done
ILLICO_CSR_PASS=2 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:fused_csr_pass2" -s 2 -c 1 -f -o $o/prof3/csr_pass2 \
    python bench.py --workload csr_ovo --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/prof3/csr_pass2.log 2>&1
