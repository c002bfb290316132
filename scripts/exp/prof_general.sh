#!/bin/bash
# ncu --set full captures of the general-path rank kernels on the data that takes them (continuous / high-count / lambda)
o=gpurun_out/prof2
mkdir -p $o
cap() {  # name workload kernel-regex skip
  local name=$1 wl=$2 re=$3 skip=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -f -o $o/$name \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/$name.log 2>&1
}
cap ovo_cont dense_ovo_continuous '^ovo_kernel' 2
cap ovr_cont dense_ovr_continuous '^ovr_kernel' 2
cap ovo_lambda dense_ovo_lambda '^ovo_kernel' 2
ls -la $o
