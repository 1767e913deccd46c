#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_postprocess.py -m gpu -q --timeout 120 -k "conv or upsample" 2>&1 | tail -n 3
for cfg in "OPB_NO_BRES=1" "OPB_NO_BRES=0"; do
  echo "== bench $cfg"
  env $cfg OPB_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/profile.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'chain ms', round(d['extra']['conv_chain_ms'],2), 'paf_up', d['extra']['paf_upsample_integrate']['frac'])"
  tail -n 44 gpurun_out/profile.txt | grep -E "conv1_1|conv1_2|conv2_1|upsample|total" | head -7
done
