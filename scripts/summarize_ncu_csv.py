#!/usr/bin/env python
"""gpurun_out/<tag>_*.raw.csv.gz (ncu --page raw --csv exports, scripts/gpu_r02j.sh) -> markdown table on stdout."""
import csv
import glob
import gzip
import os
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02j"
COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier")]
print("| capture | kernel | " + " | ".join(c[1] for c in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for f in sorted(glob.glob(os.path.join("gpurun_out", tag + "_*.raw.csv.gz"))):
    rd = list(csv.reader(gzip.open(f, "rt")))
    hdr, units, data = rd[0], rd[1], rd[2:]
    seen = set()
    for r in data:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("unnamed>::", "")
        if name in seen:
            continue
        seen.add(name)
        cells = []
        for k, _ in COLS:
            if k not in hdr:
                cells.append("-")
                continue
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            cells.append((v + " " + u).strip() if u not in ("%", "inst", "") else v)
        print("| %s | `%s` | %s |" % (os.path.basename(f).replace(".raw.csv.gz", ""), name, " | ".join(cells)))
