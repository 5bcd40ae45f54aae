"""Turn the raw ncu outputs a gpurun call brought back (gpurun_out/) into the small tracked summaries under profiles/.
    python tools/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> [<full2.ncu-rep> ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sum(a[0] for a in agg.values())} launches, {tot:.1f} us in total",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
           f"{'kernel':72s} {'launches':>8s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s}"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n:72s} {a[0]:8d} {a[1]:10.1f} {a[1] / tot:6.3f} {a[1] / a[0]:8.1f}")
    return "\n".join(out) + "\n"


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    seen, out = {}, []
    for r in rows[2:]:
        name = r[ni].split("(")[0][:70]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 2:
            continue
        rec = {"kernel": name, "instance": seen[name]}
        for k, i in idx:
            rec[k] = f"{r[i]} {units[i]}".strip()
        out.append(rec)
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w").write(launches(sys.argv[2]))
    recs = []
    for p in sys.argv[3:]:
        recs += [dict(r, source=os.path.basename(p)) for r in full(p)]
    json.dump(recs, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.json"), "w"), indent=1)
    print(f"wrote profiles/{tag}_launches.txt and profiles/{tag}_ncu_full.json ({len(recs)} kernel records)")
