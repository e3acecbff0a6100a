#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv, gzip'd) of a bench.py run -> markdown table per kernel.
usage: python scripts/summarize_launch_list.py gpurun_out/<tag>_launches_bench_c5.csv.gz [bench.json of the same command]"""
import csv
import gzip
import io
import json
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name.strip())
    name = re.sub(r"(nd::)?(\(anonymous namespace\)|<?unnamed>)::", "", name)
    return name.replace("at::native::", "torch ").strip()


lines = [l for l in gzip.open(sys.argv[1], "rt", errors="replace") if l.startswith('"')]
rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r.get("Metric Name") == "gpu__time_duration.sum"]
agg = OrderedDict()
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("%d launches, %.1f ms of kernel time in total (cold-cache, serialised: only the SHARES compare with a live run)\n" % (len(rows), tot / 1e3))
print("| kernel | launches | total ms | avg us | share |")
print("|---|---:|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if us / tot < 0.0005:
        continue
    print("| `%s` | %d | %.1f | %.1f | %.1f %% |" % (k, n, us / 1e3, us / n, 100 * us / tot))
if len(sys.argv) > 2:
    d = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][0]
    r = d["roofline"]
    print("\nLive run of the same command: `%s` %.1f us per launch (CUDA events), share of the step %.3f." % (
        r["kernel"].split(" ")[0], r["avg_launch_us"], r["share_of_step"]))
