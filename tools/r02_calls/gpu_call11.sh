#!/bin/bash
# r02 call 11: pair-wise stores for unaligned channel slices in the epilogues (libopb.so) vs the previous commit (libopb_prev.so)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P=chainer_realtime_multi-person_pose_estimation_b200
echo "=== conv tests $(date +%T)"
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/t11.log 2>&1; tail -n 3 gpurun_out/t11.log
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --precision ${PREC} --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b11_${PREC}_${name}.log 2>&1
  env "$@" OPB_PROFILE=1 timeout 600 python bench.py --precision ${PREC} --steps 3 --warmup 2 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/p11_${PREC}_${name}.log 2>&1
  python - ${PREC} ${name} <<'PY'
import json, sys, re
p, s = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/b11_%s_%s.log" % (p, s)).read().strip().splitlines()[-1])
    key = "Mconv7 " if p == "comp" else "Mconv6+7"
    c11 = [l for l in open("gpurun_out/p11_%s_%s.log" % (p, s)) if key in l][-1].split()
    print("%s %-6s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f %s %s ms sm_mhz %s" % (p, s, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], key, c11[5], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(p, s, "failed", e)
PY
}
for PREC in comp fast; do
  run new OPB_DUMMY=0
  run prev OPB_LIB_PATH=$PWD/$P/libopb_prev.so
  run new2 OPB_DUMMY=0
  run prev2 OPB_LIB_PATH=$PWD/$P/libopb_prev.so
done
echo "=== done $(date +%T)"
