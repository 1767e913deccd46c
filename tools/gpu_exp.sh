#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -s --timeout 600 -k "forward_maps or injected" 2>&1 | tail -n 6
OPB_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/profile.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'chain ms', round(d['extra']['conv_chain_ms'],2))"
tail -n 44 gpurun_out/profile.txt | grep -E "conv1_1|conv1_2|total"
