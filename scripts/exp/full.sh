#!/bin/bash
out=gpurun_out/full.log
: > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
timeout 900 python bench.py > gpurun_out/bench_default.json 2>> $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b.log 2>&1
cat $out; cat gpurun_out/bench_default.json
