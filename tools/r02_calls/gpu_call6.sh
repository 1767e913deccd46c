#!/bin/bash
# r02 call 6: comp mlp2 fusion, persistent peak kernel
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gputests6.log 2>&1; tail -n 4 gpurun_out/gputests6.log
echo "=== bench $(date +%T)"
for PREC in comp fast; do for NM in 0 1; do
  OPB_NO_MLP2=$NM timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b6_${PREC}_$NM.log 2>&1
  python - ${PREC} $NM <<'PY'
import json, sys
p, nm = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/b6_%s_%s.log" % (p, nm)).read().strip().splitlines()[-1])
    print("%s NO_MLP2=%s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 launch ms %.4f sm_mhz %s %s stage_ms %s" % (p, nm, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["extra"]["stage_ms"]))
except Exception as e:
    print(p, "failed", e)
PY
done; done
for PREC in comp fast; do
  OPB_PAIR64=1 timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b6_${PREC}_pair64.log 2>&1
  python - ${PREC} <<'PY'
import json, sys
p = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/b6_%s_pair64.log" % p).read().strip().splitlines()[-1])
    print("%s PAIR64=1 value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f sm_mhz %s %s" % (p, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print(p, "pair64 failed", e)
PY
done
echo "=== profile $(date +%T)"
OPB_PROFILE=1 timeout 600 python bench.py --precision comp --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/profile6_comp.log 2>&1; grep "opb profile" gpurun_out/profile6_comp.log | tail -n 29
echo "=== ncu $(date +%T)"
bash tools/ncu_capture.sh r02e comp smooth_nms_sep conv_mlp2 > gpurun_out/ncu_capture6.log 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|no report|inst_executed.sum|issue_active" gpurun_out/ncu_capture6.log | cut -c1-170
echo "=== done $(date +%T)"
