#!/bin/bash
o=gpurun_out
p=gpurun_out/prof_r2; mkdir -p $p
cap() {  # name workload kernel-regex skip [env...]
  local name=$1 wl=$2 re=$3 skip=$4; shift 4
  env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -f -o $p/$name \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --others none > $p/$name.log 2>&1
}
cap ovo_cont dense_ovo_continuous '^ovo_kernel' 2 A=1
cap ovr_cont dense_ovr_continuous '^ovr_kernel' 2 A=1
cap fused_wide_pass dense_ovo_lambda 'fused_wide_pass_kernel' 2 A=1
