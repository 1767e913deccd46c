#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=chainer_realtime_multi-person_pose_estimation_b200
for lib in r0e1 r0e2 r1e2; do
  echo "== lib $lib"
  OPB_LIB_PATH=$PWD/$P/libopb_$lib.so OPB_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/profile_$lib.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value']), 'conv7x7 TF', round(d['roofline']['achieved']), 'chain ms', round(d['extra']['conv_chain_ms'],2))"
  tail -n 44 gpurun_out/profile_$lib.txt | grep -E "conv1_2|conv2_1|conv3_2|Mconv1 |Mconv7x7|Mconv6|total" | head -7
done
