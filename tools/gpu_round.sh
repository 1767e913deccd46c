#!/bin/bash
# One GPU-box session: build check, probes, tests, bench.  Everything is logged to gpurun_out/.
# usage: tools/gpu_round.sh [stage ...]   stages: probe post conv kp pipe batch smoke bench benchref sanitize ncu lowres_ab
# first call of a round (parity of everything incl. the opt-in tests, bench, A/B of the experimental knobs):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh post conv kp pipe smoke bench lowres_ab'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES="${*:-probe post conv pipe smoke bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for s in $STAGES; do
  echo "=== stage $s $(date +%T)"
  case $s in
    probe) timeout 900 python tools/conv_probe.py > gpurun_out/probe.log 2>&1; tail -n 60 gpurun_out/probe.log ;;
    post)  timeout 900 python -m pytest tests/test_gpu_postprocess.py -m gpu -q --timeout 600 > gpurun_out/post.log 2>&1; tail -n 30 gpurun_out/post.log ;;
    conv)  timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 > gpurun_out/conv.log 2>&1; tail -n 30 gpurun_out/conv.log ;;
    kp)    timeout 900 python -m pytest tests/test_gpu_keypoints.py -m gpu -q -s --timeout 600 > gpurun_out/kp.log 2>&1; tail -n 40 gpurun_out/kp.log ;;
    pipe)  timeout 1500 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -s --timeout 900 > gpurun_out/pipe.log 2>&1; tail -n 40 gpurun_out/pipe.log ;;
    batch) timeout 900 python -m pytest tests/test_gpu_postprocess_batch.py -m gpu -q --timeout 600 > gpurun_out/batch.log 2>&1; tail -n 10 gpurun_out/batch.log ;;
    lowres_ab)  # the experimental low-resolution post-process variants: parity on the GPU, then bench A/B (value, e2e, stage_ms)
           OPB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_postprocess_batch.py -m gpu -q --timeout 600 > gpurun_out/lowres_parity.log 2>&1; tail -n 5 gpurun_out/lowres_parity.log
           for K in "0 0" "1 0" "2 0" "0 1" "1 1" "2 1"; do set -- $K
             OPB_FUSED_PEAKS=$1 OPB_PAF_LOWRES=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/bench_lowres_$1$2.log 2>&1
             python - "$1" "$2" <<'PY'
import json, sys
a, b = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/bench_lowres_%s%s.log" % (a, b)).read().strip().splitlines()[-1])
    print("FUSED_PEAKS=%s PAF_LOWRES=%s value %.1f e2e %.1f ms/step %.3f stage_ms %s" % (a, b, d["value"], d["e2e"]["value"], d["ms_per_step"], d["extra"]["stage_ms"]))
except Exception as e:
    print("FUSED_PEAKS=%s PAF_LOWRES=%s failed: %s" % (a, b, e))
PY
           done
           OPB_FUSED_PEAKS=2 OPB_PAF_LOWRES=1 OPB_PEAKS_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/bench_lowres_21_v2.log 2>&1
           tail -c 400 gpurun_out/bench_lowres_21_v2.log ;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 5 gpurun_out/smoke.log ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -n 3 gpurun_out/bench.log ;;
    benchref) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -n 2 gpurun_out/bench_ref.log ;;
    sanitize) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/conv_probe.py r3 > gpurun_out/sanitize.log 2>&1; echo "memcheck conv rc=$?"; tail -n 4 gpurun_out/sanitize.log; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_postprocess.py -m gpu -q -k 'synth8 or edges or upsample' >> gpurun_out/sanitize.log 2>&1; echo "memcheck post rc=$?"; tail -n 4 gpurun_out/sanitize.log ;;
    ncu)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stage-timing > gpurun_out/ncu_bench.log 2>&1; tail -n 3 gpurun_out/ncu_bench.log ;;
  esac
done
echo "=== done $(date +%T)"
