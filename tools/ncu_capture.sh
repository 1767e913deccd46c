#!/bin/bash
# One `ncu --set full` capture per hot kernel of a bench step (run under gpurun, 1 GPU), then the raw-page metrics the
# profiles/ summaries quote.  usage: tools/ncu_capture.sh <tag> <precision> [kernel-regex ...]
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture.sh r02 comp'
# Output: gpurun_out/ncu_<tag>_<precision>_<name>.ncu-rep (+ .csv of the raw page, + _summary.txt)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}; PREC=${2:-comp}; shift 2
KERNELS=("$@")
if [ ${#KERNELS[@]} -eq 0 ]; then
  KERNELS=(conv_tcgen05_swap_kernel conv_tcgen05_pair_kernel "conv_tcgen05_kernel<3, 256" conv_first smooth_nms_sep_kernel
           paf_candidates limb_assign_kernel group_persons_kernel)
fi
METRICS='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__issue_active.avg.pct|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum |l1tex__data_pipe_lsu_wavefronts_mem_shared.sum |smsp__average_warps_issue_stalled.*_per_issue_active|smsp__cycles_active.avg |sm__cycles_elapsed.max'
for K in "${KERNELS[@]}"; do
  NAME=$(echo "$K" | tr -c 'A-Za-z0-9_' '_' | sed 's/__*/_/g; s/_$//')
  OUT=gpurun_out/ncu_${TAG}_${PREC}_${NAME}
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$K" -s 2 -c 1 -f -o $OUT \
      python bench.py --precision $PREC --steps 1 --warmup 1 --no-cpu-baseline --no-parity-extra --no-stage-timing ${BENCH_EXTRA:-} > $OUT.log 2>&1
  if [ -f $OUT.ncu-rep ]; then
    ncu -i $OUT.ncu-rep --page raw --csv > $OUT.csv 2>/dev/null
    ncu -i $OUT.ncu-rep --page details --csv > ${OUT}_details.csv 2>/dev/null
    if [ "${KEEP_SOURCE:-0}" = "1" ]; then ncu -i $OUT.ncu-rep --page source --csv > ${OUT}_source.csv 2>/dev/null; fi
    # gpurun merges at most 64 MiB back: the report itself stays on the box unless asked for
    if [ "${KEEP_REP:-0}" != "1" ]; then rm -f $OUT.ncu-rep; fi
    python - "$OUT.csv" "$K" <<'PY' > ${OUT}_summary.txt
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2:]
pat = re.compile(r"gpu__time_duration.sum|dram__bytes_read.sum$|dram__bytes_write.sum$|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor.*cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__issue_active.avg.pct|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum$|smsp__average_warps_issue_stalled.*_per_issue_active|sm__cycles_elapsed.max|smsp__inst_executed.sum$|dram__bytes.sum.per_second|launch__shared_mem_per_block|launch__occupancy_limit|sm__maximum_warps_per_active_cycle_pct|achieved_occupancy")
for v in vals:
    print("# kernel:", v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else sys.argv[2])
    for h, u, x in zip(hdr, units, v):
        if pat.search(h):
            try:
                if float(x.replace(",", "")) == 0 and "stalled" in h: continue
            except ValueError: pass
            print("%-90s %16s %s" % (h, x, u))
PY
    tail -n +1 ${OUT}_summary.txt | head -60
  else
    echo "no report for $K"; tail -5 $OUT.log
  fi
done
