#!/bin/bash
# r02 evidence call: full GPU suite, per-launch event profile, ncu launch list, ncu --set full per kernel, bench lines.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gputests.log 2>&1; tail -n 6 gpurun_out/gputests.log
echo "=== smoke $(date +%T)"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
echo "=== profile $(date +%T)"
OPB_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/profile_comp.log 2>&1; grep "opb profile" gpurun_out/profile_comp.log | tail -n 36
OPB_PROFILE=1 timeout 600 python bench.py --precision fast --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/profile_fast.log 2>&1; grep "opb profile" gpurun_out/profile_fast.log | tail -n 36
echo "=== ncu launch list $(date +%T)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_comp.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/ncu_bench.log 2>&1; tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
echo "=== ncu full $(date +%T)"
KEEP_SOURCE=1 bash tools/ncu_capture.sh r02 comp conv_tcgen05_swap_kernel > gpurun_out/ncu_capture.log 2>&1
bash tools/ncu_capture.sh r02 comp conv_tcgen05_pair_kernel "conv_tcgen05_kernel<3, 256" "conv_tcgen05_kernel<3, 64, 2, 3, 9" conv_first conv_mlp2_kernel smooth_nms_sep upsample_bilinear paf_candidates limb_assign_kernel group_persons_kernel >> gpurun_out/ncu_capture.log 2>&1
grep -c "kernel:" gpurun_out/ncu_capture.log
bash tools/ncu_capture.sh r02 fast conv_tcgen05_swap_kernel >> gpurun_out/ncu_capture.log 2>&1
# the materialising PAF upsample (stage timing launches it) for the HBM half of the metric
OUT=gpurun_out/ncu_r02_comp_upsample_paf
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upsample_bilinear -c 6 -f -o $OUT python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-extra > $OUT.log 2>&1
ncu -i $OUT.ncu-rep --page raw --csv > $OUT.csv 2>/dev/null
rm -f $OUT.ncu-rep
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out | head -60
echo "=== bench $(date +%T)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_comp.log 2>&1; tail -n 1 gpurun_out/bench_comp.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
echo "=== done $(date +%T)"
