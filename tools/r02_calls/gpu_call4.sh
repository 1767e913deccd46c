#!/bin/bash
# r02 call 4: mbarrier suspend-time hint A/B, peak-kernel cell skipping, ncu of the plain-kernel layers
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P=chainer_realtime_multi-person_pose_estimation_b200
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gputests4.log 2>&1; tail -n 4 gpurun_out/gputests4.log
echo "=== A/B $(date +%T)"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/ab4_${PREC}_${name}.log 2>&1
  python - ${PREC} ${name} <<'PY'
import json, sys
p, s = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/ab4_%s_%s.log" % (p, s)).read().strip().splitlines()[-1])
    print("%s %-10s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f sm_mhz %s %s peaks %.3f" % (p, s, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["extra"]["stage_ms"]["peaks"]))
except Exception as e:
    print(p, s, "failed", e)
PY
}
for PREC in comp fast; do
  run hint20us OPB_DUMMY=0
  run hint0 OPB_LIB_PATH=$PWD/$P/libopb_hint0.so
  run hint20us_b OPB_DUMMY=0
  run hint0_b OPB_LIB_PATH=$PWD/$P/libopb_hint0.so
done
echo "=== ncu $(date +%T)"
bash tools/ncu_capture.sh r02c comp "conv_tcgen05_kernel<3, 64, 1, 4, 18" "conv_tcgen05_kernel<3, 128, 2" "conv_tcgen05_kernel<3, 256" "conv_tcgen05_kernel<1, 48" "conv_tcgen05_kernel<1, 128" smooth_nms_sep > gpurun_out/ncu_capture4.log 2>&1
bash tools/ncu_capture.sh r02c fast "conv_tcgen05_kernel<3, 64, 2, 3, 9" "conv_tcgen05_kernel<3, 128, 2" conv_mlp2 >> gpurun_out/ncu_capture4.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|no report" gpurun_out/ncu_capture4.log | cut -c1-150
echo "=== done $(date +%T)"
