#!/bin/bash
# Round profiles: bench lines of the four K562 workloads (fused paths, and the general two-kernel paths with the fused
# ones switched off) + reference arm, launch list of the default bench command, one `ncu --set full` capture per kernel.
# Outputs under gpurun_out/prof/ (summarised into profiles/ afterwards).
o=gpurun_out/prof
mkdir -p $o
for wl in dense_ovo dense_ovr csr_ovo csr_ovr; do
  timeout 900 python bench.py --workload $wl > $o/bench_$wl.json 2> $o/bench_$wl.err
  ILLICO_OVO_FUSED=0 ILLICO_OVR_FUSED=0 timeout 900 python bench.py --workload $wl --no-e2e --no-cpu-baseline > $o/bench_general_$wl.json 2> $o/bench_general_$wl.err
done
for wl in dense_ovo dense_ovr csr_ovo csr_ovr; do
  timeout 900 python bench.py --workload $wl --continuous --no-e2e --no-cpu-baseline --steps 3 > $o/bench_continuous_$wl.json 2> $o/bench_continuous_$wl.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_reference.json 2> $o/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_dense_ovo.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $o/launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_csr_ovo.csv \
  python bench.py --workload csr_ovo --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $o/launches_csr.log 2>&1
cap() {  # name workload kernel-regex [env...]
  local name=$1 wl=$2 re=$3; shift 3
  env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$re" -c 1 -f -o $o/$name \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $o/$name.log 2>&1
}
cap fused_pass dense_ovo fused_pass_kernel A=1
cap fused_pass_ovr dense_ovr fused_pass_kernel A=1
cap fused_epilogue dense_ovo fused_epilogue_kernel A=1
cap fused_csr_pass csr_ovo fused_csr_pass_kernel A=1
cap stage_tma dense_ovr stage_dense_tma_kernel ILLICO_OVR_FUSED=0
cap ovr dense_ovr 'ovr_.*kernel' ILLICO_OVR_FUSED=0
cap ovo csr_ovo 'illico::ovo_kernel' ILLICO_OVO_FUSED=0
cap stage_csr csr_ovo stage_csr_kernel ILLICO_OVO_FUSED=0
ls -la $o | head -50
