#!/bin/bash
# r02 final evidence: smoke, bench lines (comp default / fast / reference arm), ncu launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== smoke $(date +%T)"
timeout 600 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -n 2 gpurun_out/final_smoke.log
echo "=== bench comp $(date +%T)"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_comp.log 2>&1; tail -n 1 gpurun_out/final_bench_comp.log | cut -c1-300
echo "=== bench fast $(date +%T)"
timeout 900 python bench.py --precision fast --steps 10 --warmup 3 --no-cpu-baseline --no-parity-extra > gpurun_out/final_bench_fast.log 2>&1; tail -n 1 gpurun_out/final_bench_fast.log | cut -c1-300
echo "=== bench reference $(date +%T)"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; tail -n 1 gpurun_out/final_bench_ref.log | cut -c1-300
echo "=== ncu launch list $(date +%T)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/final_launches_comp.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stage-timing --no-parity-extra > gpurun_out/final_ncu_bench.log 2>&1; tail -n 1 gpurun_out/final_ncu_bench.log | cut -c1-200
echo "=== done $(date +%T)"
