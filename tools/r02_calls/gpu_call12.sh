#!/bin/bash
# r02 call 12: ncu of the kernels changed last (conv1_1 with halo prefetch, 1x1 heads / fused pair with pair-wise stores)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
bash tools/ncu_capture.sh r02h comp conv_first 'conv_tcgen05_kernel<\(int\)1, \(int\)48' > gpurun_out/ncu_capture12.log 2>&1
bash tools/ncu_capture.sh r02h fast conv_first conv_mlp2 >> gpurun_out/ncu_capture12.log 2>&1
grep -E "kernel:|time_duration|dram__bytes.sum.per_second|issue_active|no report" gpurun_out/ncu_capture12.log | cut -c1-170
