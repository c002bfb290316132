#!/bin/bash
# Round profiles: bench lines of the four K562 workloads + reference arm, launch list of the default bench command,
# one `ncu --set full` capture per kernel.  Outputs under gpurun_out/prof/ (summarised into profiles/ afterwards).
o=gpurun_out/prof
mkdir -p $o
for wl in dense_ovo dense_ovr csr_ovo csr_ovr; do
  timeout 900 python bench.py --workload $wl > $o/bench_$wl.json 2> $o/bench_$wl.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_reference.json 2> $o/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_dense_ovo.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $o/launches.log 2>&1
cap() {  # name workload kernel-regex
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$3" -c 1 -f -o $o/$1 \
    python bench.py --workload $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $o/$1.log 2>&1
}
cap fused dense_ovo fused_pass_kernel
cap stage_tma dense_ovr stage_dense_tma_kernel
cap ovr dense_ovr 'ovr_.*kernel'
cap ovo csr_ovo '^ovo_kernel|illico::ovo_kernel'
cap stage_csr csr_ovo stage_csr_kernel
ls -la $o
