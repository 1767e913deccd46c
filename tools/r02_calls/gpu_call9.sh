#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== new tests $(date +%T)"
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "scheduling_variants" > gpurun_out/t9.log 2>&1; tail -n 3 gpurun_out/t9.log
echo "=== drain_seg A/B $(date +%T)"
for S in 4 7 4 7; do
  OPB_DRAIN_SEG=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/b9_$S.log 2>&1
  python - $S <<'PY'
import json, sys
s = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/b9_%s.log" % s).read().strip().splitlines()[-1])
    print("DRAIN_SEG=%s value %.1f e2e %.1f ms/step %.3f conv_chain_ms %.3f 7x7 %.4f sm_mhz %s" % (s, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["conv_chain_ms"], d["roofline"]["ms_per_launch"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(s, "failed", e)
PY
done
echo "=== done $(date +%T)"
