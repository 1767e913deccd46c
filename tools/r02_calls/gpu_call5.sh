#!/bin/bash
# r02 call 5: cell-based peak kernel; ncu of the plain-kernel layers (demangled template names)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gputests5.log 2>&1; tail -n 4 gpurun_out/gputests5.log
echo "=== bench $(date +%T)"
for PREC in comp fast; do
  timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b5_${PREC}.log 2>&1
  python - ${PREC} <<'PY'
import json, sys
p = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/b5_%s.log" % p).read().strip().splitlines()[-1])
    print("%s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f sm_mhz %s %s stage_ms %s" % (p, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["extra"]["stage_ms"]))
except Exception as e:
    print(p, "failed", e)
PY
done
echo "=== ncu $(date +%T)"
bash tools/ncu_capture.sh r02d comp smooth_nms_sep 'conv_tcgen05_kernel<\(int\)3, \(int\)64, \(int\)1' 'conv_tcgen05_kernel<\(int\)3, \(int\)128, \(int\)2' 'conv_tcgen05_kernel<\(int\)3, \(int\)256' 'conv_tcgen05_kernel<\(int\)1, \(int\)48' 'conv_tcgen05_kernel<\(int\)1, \(int\)128' > gpurun_out/ncu_capture5.log 2>&1
bash tools/ncu_capture.sh r02d fast 'conv_tcgen05_kernel<\(int\)3, \(int\)64, \(int\)2' 'conv_tcgen05_kernel<\(int\)3, \(int\)128, \(int\)2' 'conv_tcgen05_kernel<\(int\)3, \(int\)256' >> gpurun_out/ncu_capture5.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|no report" gpurun_out/ncu_capture5.log | cut -c1-170
echo "=== done $(date +%T)"
