#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ (run here, no GPU needed).
  python profiles/summarize_ncu.py launches <launches.csv>      per-kernel launch count / total / average / share
  python profiles/summarize_ncu.py full <report.ncu-rep>        per-launch DRAM bytes, DRAM %, occupancy, L2 hit, ...
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg[r[ki].split("(")[0]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'total us':>11s} {'avg us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {v[0]:5d} {v[1]:11.1f} {v[1] / v[0]:9.1f} {v[1] / tot:6.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd MB"),
            ("dram__bytes_write.sum", "wr MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
            ("lts__t_sector_hit_rate.pct", "L2hit%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%")]
    idx = [hdr.index(w) if w in hdr else -1 for w, _ in want]
    print(" | ".join(n for _, n in want))
    for r in rows[2:]:
        vals = []
        for (w, n), i in zip(want, idx):
            x = r[i] if i >= 0 else "NA"
            if n == "kernel":
                x = x.split("(")[0][:28]
            else:
                try:
                    x = f"{float(x.replace(',', '')):.1f}"
                except ValueError:
                    pass
            vals.append(x)
        print(" | ".join(vals))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
