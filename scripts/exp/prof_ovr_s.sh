#!/bin/bash
o=gpurun_out/prof3; mkdir -p $o
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^ovr_kernel" -s 2 -c 1 -f -o $o/ovr_u \
    python bench.py --workload dense_ovr_unique --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/ovr_u.log 2>&1
ls -la $o/ovr_u.ncu-rep
