#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests (post incl. new ones)"
timeout 600 python -m pytest tests/test_gpu_postprocess.py -m gpu -q --timeout 120 -k "candidate or capacity" 2>&1 | tail -n 3
echo "== conv tests with OPB_SWAP=3"
OPB_SWAP=3 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 -k "fast" 2>&1 | tail -n 12
for cfg in "OPB_SWAP=0" "OPB_SWAP=1" "OPB_SWAP=3"; do
  echo "== bench $cfg"
  env $cfg OPB_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/profile_x.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value']), 'conv7x7 TF', round(d['roofline']['achieved']), 'chain ms', round(d['extra']['conv_chain_ms'],2))"
  tail -n 44 gpurun_out/profile_x.txt | grep -E "Mconv1 |Mconv7x7|conv5_1|conv2_2|total" | head -5
done
