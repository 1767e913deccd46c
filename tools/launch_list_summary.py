#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/rNN_ncu_launch_list_summary.txt).
usage: python tools/launch_list_summary.py gpurun_out/final_launches_comp.csv [bench.json]"""
import csv
import json
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*$", "", r[ik])[:96]
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(r[iu], 1e-6)
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-98s n=%4d total %9.3f ms  %5.1f%%" % (name, n, t, 100.0 * t / total))
    print("total %.3f ms over %d launches" % (total, sum(n for n, _ in tot.values())))
    dom = [(k, v) for k, v in tot.items() if "swap7" in k]
    if dom:
        n, t = sum(v[0] for _, v in dom), sum(v[1] for _, v in dom)
        print("\ndominant kernel (conv_tcgen05_swap7_kernel, 20 launches per pass) share of a pass:")
        print("  ncu launch list %.1f%% (%d launches)" % (100.0 * t / total, n))
        if len(sys.argv) > 2:
            d = json.loads([l for l in open(sys.argv[2]).read().splitlines() if l.startswith("{")][-1])
            ms = d["roofline"]["ms_per_launch"]
            sync = d["value_api"]["sync_api"]["ms_per_step"]
            print("  live CUDA-event timing in bench.py: 20 x %.4f ms = %.2f ms = %.1f%% of the %.2f ms one-batch-at-a-time step"
                  % (ms, 20 * ms, 100.0 * 20 * ms / sync, sync))
            print("  (%.1f%% of the %.2f ms per batch with two batches in flight, where kernels of the two slots overlap)"
                  % (100.0 * 20 * ms / d["ms_per_step"], d["ms_per_step"]))


if __name__ == "__main__":
    main()
