#!/bin/bash
# r02 call 3: restructured plain-kernel issuer; A/B of CTA pairs on the 3x3 layers; per-launch profiles
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gputests3.log 2>&1; tail -n 4 gpurun_out/gputests3.log
echo "=== A/B $(date +%T)"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/ab3_${PREC}_${name}.log 2>&1
  python - ${PREC} ${name} <<'PY'
import json, sys
p, s = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/ab3_%s_%s.log" % (p, s)).read().strip().splitlines()[-1])
    print("%s %-10s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f sm_mhz %s %s" % (p, s, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print(p, s, "failed", e)
PY
}
for PREC in comp fast; do
  run default OPB_DUMMY=0
  run pair3 OPB_PAIR=3
  run pair7 OPB_PAIR=7
done
echo "=== profile $(date +%T)"
for PREC in comp fast; do
OPB_PROFILE=1 timeout 600 python bench.py --precision $PREC --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/profile3_$PREC.log 2>&1; grep "opb profile" gpurun_out/profile3_$PREC.log | tail -n 30 | head -24
done
echo "=== ncu $(date +%T)"
bash tools/ncu_capture.sh r02c comp "conv_tcgen05_kernel<3, 64, 1, 4, 18" "conv_tcgen05_kernel<3, 128, 2" "conv_tcgen05_kernel<3, 256" conv_first conv_mlp2 > gpurun_out/ncu_capture3.log 2>&1
bash tools/ncu_capture.sh r02c fast "conv_tcgen05_kernel<3, 64, 2, 3, 9" "conv_tcgen05_kernel<3, 128, 2" >> gpurun_out/ncu_capture3.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|no report" gpurun_out/ncu_capture3.log | cut -c1-150
echo "=== done $(date +%T)"
