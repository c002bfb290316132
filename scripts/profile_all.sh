#!/bin/bash
# Round profiles (one gpurun call, one box): the default bench line (with every other workload), the reference arm,
# the launch list of the default bench command, one `ncu --set full` capture per kernel on the workload that exercises it.
# Outputs under gpurun_out/prof_r2/ (summarised into profiles/ afterwards by scripts/ncu_summary.py).
o=gpurun_out/prof_r2
mkdir -p $o
timeout 1500 python bench.py > $o/bench_default.json 2> $o/bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_reference.json 2> $o/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_dense_ovo.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_csr_ovo.csv \
  python bench.py --workload csr_ovo --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/launches_csr.log 2>&1
cap() {  # name workload kernel-regex skip [env...]
  local name=$1 wl=$2 re=$3 skip=$4; shift 4
  env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -f -o $o/$name \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --others none > $o/$name.log 2>&1
}
cap fused_pass dense_ovo 'fused_pass_kernel' 2 A=1
cap fused_pass_ovr dense_ovr 'fused_pass_kernel' 2 A=1
cap fused_epilogue dense_ovo 'fused_epilogue_kernel' 2 A=1
cap fused_wide_pass dense_ovo_lambda 'fused_wide_pass_kernel' 2 A=1
cap fused_csr_pass csr_ovo 'fused_csr_pass_kernel' 2 A=1
cap stage_tma dense_ovo_continuous 'stage_dense_tma_kernel' 5 A=1
cap ovo_cont dense_ovo_continuous '^void illico::ovo_kernel|illico::ovo_kernel' 2 A=1
cap ovr_cont dense_ovr_continuous 'illico::ovr_kernel' 2 A=1
cap ovr_table dense_ovr 'ovr_table_kernel' 2 ILLICO_OVR_FUSED=0
cap stage_csr csr_ovo 'stage_csr_kernel' 2 ILLICO_OVO_FUSED=0
cap stage_csc backed_csc_ovr 'stage_csc_kernel' 2 A=1
ls -la $o | head -60
