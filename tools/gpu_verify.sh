#!/bin/bash
# final verification of the committed tree: the driver's own sequence (GPU tests with -x, smoke, default bench, reference arm)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/verify_tests.log 2>&1; tail -n 3 gpurun_out/verify_tests.log
echo "=== smoke $(date +%T)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; tail -n 1 gpurun_out/verify_smoke.log
echo "=== bench (defaults) $(date +%T)"
timeout 1200 python bench.py > gpurun_out/verify_bench.log 2>&1; tail -n 1 gpurun_out/verify_bench.log | cut -c1-250
python - <<'PY'
import json
d = json.loads(open("gpurun_out/verify_bench.log").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms/step %.3f roofline.frac %.3f issued %.0f clocks %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["issued_tensor_work"]["fp16_equivalent_tflops"], d["clocks"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None))
PY
echo "=== done $(date +%T)"
