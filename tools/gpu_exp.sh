#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_postprocess.py -m gpu -q --timeout 300 2>&1 | tail -n 3
OPB_PROFILE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/profile.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'chain ms', round(d['extra']['conv_chain_ms'],2), d['extra']['stage_ms'])"
tail -n 44 gpurun_out/profile.txt | grep -E "tile_max|smooth|total"
