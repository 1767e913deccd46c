#!/bin/bash
# r02 call 8: pair kernel unit mode (Mconv1 computes 88 instead of 96 columns)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gputests8.log 2>&1; tail -n 4 gpurun_out/gputests8.log
echo "=== bench $(date +%T)"
for PREC in comp fast; do for L in 1 0 1 0; do
  OPB_PAIR_UNITS=$L timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b8_${PREC}_$L.log 2>&1
  python - ${PREC} $L <<'PY'
import json, sys
p, nm = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/b8_%s_%s.log" % (p, nm)).read().strip().splitlines()[-1])
    print("%s PAIR_UNITS=%s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f sm_mhz %s %s" % (p, nm, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print(p, "failed", e)
PY
done; done
echo "=== profile $(date +%T)"
for L in 1 0; do
OPB_PAIR_UNITS=$L OPB_PROFILE=1 timeout 600 python bench.py --precision comp --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/profile8_comp_$L.log 2>&1; grep "opb profile" gpurun_out/profile8_comp_$L.log | grep -E "Mconv1|total" | tail -n 2
done
echo "=== ncu $(date +%T)"
bash tools/ncu_capture.sh r02g comp conv_tcgen05_pair_kernel > gpurun_out/ncu_capture8.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active" gpurun_out/ncu_capture8.log | cut -c1-170
echo "=== done $(date +%T)"
