#!/bin/bash
# N GPUs of one box: single-process devices= against one GPU, then bench.py under torchrun (weak line + strong block)
N=${1:-2}
o=gpurun_out
{
echo "== scripts/exp/multi_gpu_check.py (asymptotic_wilcoxon(devices=...) against the single-GPU call), $N GPUs"
python scripts/exp/multi_gpu_check.py 2>&1 | grep -v Warning
echo "== bench.py --gpus $N (torchrun, one rank per GPU)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -E '^\{'
if [ "$N" = "8" ]; then
echo "== bench.py --workload c5_shard --gpus 8 (configs[4]: 2M cells x 20000 genes CSR, 10001 groups, 2500 genes per GPU)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5_shard --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{'
fi
} > $o/multi_gpu_$N.txt 2>&1
tail -c 1500 $o/multi_gpu_$N.txt
