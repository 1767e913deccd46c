#!/bin/bash
# r02 call 2: elect.sync issue path + lean 7x7 swap kernel: GPU parity, A/B, per-launch profile, ncu of the new kernel
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gputests2.log 2>&1; tail -n 4 gpurun_out/gputests2.log
echo "=== A/B $(date +%T)"
for PREC in fast comp; do for S7 in 0 1; do
  OPB_SWAP7=$S7 timeout 600 python bench.py --precision $PREC --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/ab_${PREC}_swap7_$S7.log 2>&1
  python - $PREC $S7 <<'PY'
import json, sys
p, s = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/ab_%s_swap7_%s.log" % (p, s)).read().strip().splitlines()[-1])
    print("%s SWAP7=%s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f clocks %s" % (p, s, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]))
except Exception as e:
    print(p, s, "failed", e)
PY
done; done
echo "=== profile $(date +%T)"
for PREC in comp fast; do
OPB_PROFILE=1 timeout 600 python bench.py --precision $PREC --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/profile2_$PREC.log 2>&1; grep "opb profile" gpurun_out/profile2_$PREC.log | tail -n 30
done
echo "=== ncu $(date +%T)"
KEEP_SOURCE=1 bash tools/ncu_capture.sh r02b comp conv_tcgen05_swap7_kernel > gpurun_out/ncu_capture2.log 2>&1
bash tools/ncu_capture.sh r02b fast conv_tcgen05_swap7_kernel >> gpurun_out/ncu_capture2.log 2>&1
KEEP_SOURCE=1 bash tools/ncu_capture.sh r02b comp "conv_tcgen05_kernel<3, 64, 2, 3, 9" conv_tcgen05_pair_kernel >> gpurun_out/ncu_capture2.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active" gpurun_out/ncu_capture2.log | cut -c1-150
echo "=== done $(date +%T)"
