#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
It executes oracle/_ref/{PHOLD,TrafficSim} (reference sources compiled in place against the stand-in
headers of oracle/shim, single MPI rank) and stores, per case, the sorted committed lines (small cases) or
their count + sha256 + order-independent digest (large cases). The ring inputs are stored as arrays so that the
tests do not need the reference tree on the GPU box. Known answers quoted by the reference itself
(traffic/README.md:119-126) are stored next to the generated ones.
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from scalesim_b200 import capi, traffic  # noqa: E402


def run_ref(binary, args, threads=1, retries=8, cwd=None):
    """The reference crashes at start-up now and then (unlocked stopwatch registry, stopwatch.hpp:50-56): retry."""
    env = dict(os.environ, SCALESIM_NUM_THR=str(threads))
    for _ in range(retries):
        p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", binary)] + [str(a) for a in args], env=env,
                           cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if p.returncode == 0:
            lines = sorted(l for l in p.stdout.decode().splitlines() if l.startswith("RE,"))
            return lines, p.stderr.decode()
    raise RuntimeError(f"{binary} kept failing: rc={p.returncode}")


def lines_to_records(lines):
    rec = np.zeros(len(lines), dtype=capi.RECORD_DTYPE)
    for i, l in enumerate(lines):
        f = l.split(",")
        rec[i] = (int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]),
                  {"START": 1, "END": 2}.get(f[6], 0) if len(f) > 6 else 0)
    return rec


def summary(lines):
    blob = "".join(l + "\n" for l in lines).encode()
    d = capi.digest_records(lines_to_records(lines))
    return {"count": len(lines), "sha256": hashlib.sha256(blob).hexdigest(), "digest": [str(x) for x in d]}


def stat(stderr, label):
    for l in stderr.splitlines():
        if label in l:
            return int(l.split(":")[-1].strip().split()[0])
    return None


def main():
    # ---- ring (BASELINE.json config 1) ----
    ring = os.path.join(REF, "traffic", "ring")
    sc = traffic.read_scenario(os.path.join(ring, "rd.sim.csv"), os.path.join(ring, "trip.csv"),
                               os.path.join(ring, "part", "graph.part.4"))
    with tempfile.TemporaryDirectory() as td:
        part0 = os.path.join(td, "part0")
        with open(part0, "w") as f:
            f.write("0\n" * 10)            # single-rank harness: every LP in partition 0
        lines, err = run_ref("TrafficSim", [part0, os.path.join(ring, "rd.sim.csv"), os.path.join(ring, "trip.csv")],
                             threads=3)
    out = {
        "source": "oracle/_ref/TrafficSim on traffic/ring/{rd.sim.csv,trip.csv}, finish 10800",
        "n_lp": int(sc.n_lp),
        "state_ids": sc.state_ids.tolist(), "road_start": sc.road_start.tolist(), "road_dst": sc.road_dst.tolist(),
        "road_speed": sc.road_speed.tolist(), "road_len": sc.road_len.tolist(),
        "trip_id": sc.trip_id.tolist(), "trip_dep": sc.trip_dep.tolist(), "route_start": sc.route_start.tolist(),
        "route_nodes": sc.route_nodes.tolist(), "partition4": sc.partition.tolist(),
        "lines": lines,
        "readme_lines": ["RE,0,2,83,3,124", "RE,1,3,92,4,133", "RE,2,2,94,3,135", "RE,0,3,124,4,165",
                         "RE,1,4,133,5,174", "RE,2,3,135,4,176"],   # traffic/README.md:121-126
        "processed": stat(err, "# of processing events"),
    }
    out.update(summary(lines))
    json.dump(out, open(os.path.join(HERE, "ring.json"), "w"), indent=1)
    print("ring", out["count"], out["sha256"])

    # ---- PHOLD small: full sorted lines, 1 and 4 worker threads (rollbacks) must agree ----
    l1, e1 = run_ref("PHOLD", [100, 1600, 0.1, 1.0, 1, 50], threads=1)
    l4, e4 = run_ref("PHOLD", [100, 1600, 0.1, 1.0, 1, 50], threads=4)
    assert l1 == l4, "reference output depends on thread count?!"
    out = {"source": "oracle/_ref/PHOLD 100 1600 0.1 1.0 1 50 (1 and 4 worker threads, identical)",
           "args": [100, 1600, 0.1, 1.0],
           "cancels_1thr": stat(e1, "# of cancels"), "cancels_4thr": stat(e4, "# of cancels")}
    out.update(summary(l1))
    json.dump(out, open(os.path.join(HERE, "phold_100_1600.json"), "w"), indent=1)
    with gzip.open(os.path.join(HERE, "phold_100_1600.lines.gz"), "wt") as f:
        f.write("".join(l + "\n" for l in l1))
    print("phold 100", out["count"], out["sha256"], "cancels", out["cancels_1thr"], out["cancels_4thr"])

    # ---- PHOLD other small shapes (remote ratio 0.5, lambda 2) ----
    for args in ([64, 1024, 0.5, 1.0], [1000, 4000, 0.25, 2.0]):
        l, e = run_ref("PHOLD", args + [1, 50], threads=2)
        out = {"source": f"oracle/_ref/PHOLD {' '.join(map(str, args))} 1 50 (2 worker threads)", "args": args,
               "cancels": stat(e, "# of cancels")}
        out.update(summary(l))
        json.dump(out, open(os.path.join(HERE, f"phold_{args[0]}_{args[1]}.json"), "w"), indent=1)
        print("phold", args, out["count"], out["sha256"])

    # ---- PHOLD medium: count + sha256 + digest only ----
    l, e = run_ref("PHOLD", [10000, 160000, 0.1, 1.0, 1, 50], threads=4)
    out = {"source": "oracle/_ref/PHOLD 10000 160000 0.1 1.0 1 50 (4 worker threads)", "args": [10000, 160000, 0.1, 1.0],
           "cancels": stat(e, "# of cancels"), "sim_ms": stat(e, "Simulation elapsed time")}
    out.update(summary(l))
    json.dump(out, open(os.path.join(HERE, "phold_10000_160000.json"), "w"), indent=1)
    print("phold 10000", out["count"], out["sha256"], "ms", out["sim_ms"])


if __name__ == "__main__":
    main()
