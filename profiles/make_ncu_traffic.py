#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report (run here, no GPU needed):
    python profiles/make_ncu_traffic.py gpurun_out/<report>.ncu-rep "<what was captured>"
Per kernel: DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches
that did work (launches shorter than 8 us — empty routes — are left out). bench.py reads the dominant kernel's figure
into roofline.traffic."""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(path, what):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, it = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[2:]:
        us = float(r[it].replace(",", "")) * tscale.get(units[it], 1.0)
        if us < 8.0:
            continue
        name = r[ik].split("(")[0].replace("void ", "").split("<")[0].replace("ssb::", "")
        b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        a = agg[name]
        a[0] += 1; a[1] += b; a[2] += us
    res = {"_source": what}
    for k, (n, b, us) in agg.items():
        res[k] = {"dram_bytes_per_launch": b / n, "launches_captured": n, "avg_us_under_ncu": us / n}
    json.dump(res, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
