#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== profile"; OPB_PROFILE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/profile.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['achieved'], d['extra'])"
tail -n 47 gpurun_out/profile.txt | head -36
