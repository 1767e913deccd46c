#!/bin/bash
# per-launch CUDA-event profile of the default fast path (OPB_PROFILE=1), batch 32, one pipeline pass per step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OPB_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing > gpurun_out/profile.txt 2>&1
grep -c "opb profile" gpurun_out/profile.txt
