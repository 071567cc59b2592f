#!/bin/bash
# Sharded configurations on N GPUs of one box (run under `gpurun --gpus N`):  tools/multi_gpu_round.sh N [tag]
# ensemble10 with the three exchange routes, brats50, and (N = 2 only) the NCCL / peer-memory tests.
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_n${N}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29600
for ex in peer nccl torch; do
  PORT=$((PORT + 1))
  timeout 300 $TR --master-port $PORT bench.py --gpus $N --config ensemble10 --exchange $ex --steps 5 --warmup 3 > gpurun_out/${TAG}_ens10_n${N}_$ex.log 2>&1
  echo "ensemble10 $ex rc=$?"; grep '^{' gpurun_out/${TAG}_ens10_n${N}_$ex.log | tail -1 | cut -c1-400
done
PORT=$((PORT + 1))
timeout 600 $TR --master-port $PORT bench.py --gpus $N --config brats50 --steps 2 --warmup 1 > gpurun_out/${TAG}_brats50_n${N}.log 2>&1
echo "brats50 rc=$?"; grep '^{' gpurun_out/${TAG}_brats50_n${N}.log | tail -1 | cut -c1-400
