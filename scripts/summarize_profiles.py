"""Turn what a gpurun evidence call brought back (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag>"""
import csv, gzip, io, json, os, re, shutil, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name.strip())
    return re.sub(r"^(nd::)?(\(anonymous namespace\)|<?unnamed>)::", "", name).strip()

# ---- launch list
lp = os.path.join(G, tag + "_launches_l2t.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    # profile_step runs identical steps; a step starts at the first encoder kernel after a generator launch.
    # Keep the LAST COMPLETE step.
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    starts = [i for i, r in enumerate(rows) if i == 0 or ("generator" in rows[i - 1]["Kernel Name"] and "embed" not in r["Kernel Name"]
                                                          and "generator" not in r["Kernel Name"])]
    agg = OrderedDict()
    step = rows[starts[-2]:starts[-1]] if len(starts) >= 2 else rows
    for r in step:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        key = (short(r["Kernel Name"]), r.get("Grid Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, tag + "_launch_summary_l2t.md"), "w") as f:
        f.write("# %s — ncu launch list, last complete step of the bench workload (L2T greedy, B=1024, T=512, L=100)\n\n" % tag)
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv python scripts/profile_step.py l2t 1`\n"
                "(per-launch times are cold-cache and serialised: compare SHARES with the event-timed profile, not absolutes).\n"
                "Raw list: `%s_launches_l2t.csv.gz`.\n\n| kernel | grid | launches | total ms | avg us | share |\n|---|---|---:|---:|---:|---:|\n" % tag)
        for (k, g), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %s | %d | %.2f | %.1f | %.1f %% |\n" % (k, g, n, us / 1e3, us / n, 100 * us / tot))
        f.write("| **total** | | %d | %.2f | | |\n" % (sum(a[0] for a in agg.values()), tot / 1e3))
    with open(lp, "rb") as fi, gzip.open(os.path.join(P, tag + "_launches_l2t.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)

# ---- full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k in ("cross_attn", "lstm", "gemm", "self_attn", "generator", "cross_ring"):
    rep = os.path.join(G, "%s_prof_%s.ncu-rep" % (tag, k))
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with gzip.open(os.path.join(P, "%s_ncu_%s_raw.csv.gz" % (tag, k)), "wt") as f:
        f.write(raw)
    with open(os.path.join(P, "%s_ncu_%s.md" % (tag, k)), "w") as f:
        f.write("# %s — `ncu --set full --clock-control none --import-source on -k regex:%s` (scripts/profile_step.py l2t 1)\n\n" % (tag, k))
        f.write("Full raw page: `%s_ncu_%s_raw.csv.gz`.\n\n" % (tag, k))
        for r in data:
            f.write("## `%s` grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for w in WANT:
                if w in hdr:
                    f.write("| %s | %s | %s |\n" % (w, r[hdr.index(w)], units[hdr.index(w)]))
            f.write("\n")
    if k == "cross_attn" and data:
        def val(r, w):
            v, u = float(r[hdr.index(w)].replace(",", "")), units[hdr.index(w)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        tr = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in data]
        json.dump({"kernel": short(data[0][hdr.index("Kernel Name")]), "dram_bytes_per_launch": int(sum(tr) / len(tr)),
                   "launches": len(tr), "source": "%s_ncu_cross_attn_raw.csv.gz (dram__bytes_read.sum + dram__bytes_write.sum)" % tag},
                  open(os.path.join(P, "cross_attn_traffic.json"), "w"), indent=1)

# ---- event profiles, bench lines, test log
for fn in sorted(os.listdir(G)):
    if fn.startswith(tag + "_profile_") or fn in (tag + "_bench.json", tag + "_bench_ref.json", tag + "_summary.txt"):
        shutil.copy(os.path.join(G, fn), os.path.join(P, fn))
print("profiles/ updated for", tag)
