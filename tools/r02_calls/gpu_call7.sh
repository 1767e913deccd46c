#!/bin/bash
# r02 call 7: LPT tile schedule of the 7x7 kernel
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gputests7.log 2>&1; tail -n 4 gpurun_out/gputests7.log
echo "=== bench $(date +%T)"
for PREC in comp fast; do for L in 1 0 1 0; do
  OPB_SWAP7_LPT=$L timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b7_${PREC}_$L.log 2>&1
  python - ${PREC} $L <<'PY'
import json, sys
p, nm = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/b7_%s_%s.log" % (p, nm)).read().strip().splitlines()[-1])
    print("%s LPT=%s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f sm_mhz %s %s" % (p, nm, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print(p, "failed", e)
PY
done; done
echo "=== ncu $(date +%T)"
bash tools/ncu_capture.sh r02f comp conv_tcgen05_swap7_kernel > gpurun_out/ncu_capture7.log 2>&1
bash tools/ncu_capture.sh r02f fast conv_tcgen05_swap7_kernel >> gpurun_out/ncu_capture7.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct|cycles_active.avg|cycles_elapsed.max " gpurun_out/ncu_capture7.log | cut -c1-170
echo "=== done $(date +%T)"
