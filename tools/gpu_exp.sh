#!/bin/bash
# full ncu capture of the dominant kernel (grouped 7x7 128->128) + 2-GPU bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${1:-}" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel -s 40 -c 2 -o gpurun_out/prof_conv7x7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -n 3 gpurun_out/ncu_full.log | cut -c 1-300
  ls -la gpurun_out/*.ncu-rep
fi
if [ "${1:-}" = "multi" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -n 3 | cut -c 1-900
fi
