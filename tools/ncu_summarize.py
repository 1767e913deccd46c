#!/usr/bin/env python
"""profiles/ summary of one `ncu --set full` raw-page csv (tools/ncu_capture.sh): the metrics DESIGN.md and bench.py quote.
usage: python tools/ncu_summarize.py gpurun_out/ncu_<...>.csv "<header text>" > profiles/<name>_summary.txt"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    for v in rows[2:]:
        name = v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("\nKernel Name".ljust(92) + name[:160])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-91s %s %s" % (k, v[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    x = float(v[i].replace(",", ""))
                except ValueError:
                    continue
                if x >= 0.05:
                    stalls.append((x, h[len(STALL):-len("_per_issue_active.ratio")]))
        print("stall cycles per issued instruction (>= 0.05): " + ", ".join("%s %.2f" % (n, x) for x, n in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main()
