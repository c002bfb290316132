#!/bin/bash
# ncu --set full capture of the wide fused pass on the lambda workload
o=gpurun_out/prof3
mkdir -p $o
cap() {  # name workload kernel-regex skip
  local name=$1 wl=$2 re=$3 skip=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -f -o $o/$name \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/$name.log 2>&1
}
cap wide_lambda dense_ovo_lambda '^fused_wide_pass' 2
ls -la $o
