#!/usr/bin/env python
"""Golden fixture for the state-change (SC) what-if query from the UNMODIFIED reference (build container only):
    python tests/golden/make_golden_sc.py
Runs oracle/_ref/TrafficSim --diff_init on the ring example and then --diff_repeat with traffic/ring/change_state.csv
at 2 / 3 / 4 / 8 worker threads, several times, and adds to ring_diff.json:
  sc_lines      the query file
  sc_variants   every distinct output (sorted lines) with the thread counts that produced it
  sc_new_lines  the lines of the repeat output that the init run did not have — identical in every variant
The reference's SC output is NOT one set: besides the new trajectories it keeps stale ones, because an anti-message
that reaches an LP before that LP has lazily loaded its history is dropped (logical_process.hpp:132-138 appends the
loaded events behind it, queue.hpp:87-90 erases nothing) — which anti-messages arrive early depends on the host
threads (69 lines at 2 threads, 70 at 3 / 4 / 8 in every run so far)."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
from make_golden import run_ref  # noqa: E402


def main():
    ring = os.path.join(REF, "traffic", "ring")
    sc_file = os.path.join(ring, "change_state.csv")
    variants = {}
    with tempfile.TemporaryDirectory() as td:
        part0 = os.path.join(td, "part0")
        with open(part0, "w") as f:
            f.write("0\n" * 10)
        st = os.path.join(td, "st")
        os.makedirs(st)
        init, _ = run_ref("TrafficSim", ["--sim_id=7", f"--store_dir={st}", "--diff_init", part0,
                                         os.path.join(ring, "rd.sim.csv"), os.path.join(ring, "trip.csv")], threads=2)
        for thr in (2, 3, 4, 8):
            for _ in range(3):
                lines, _ = run_ref("TrafficSim", ["--sim_id=7", f"--store_dir={st}", "--diff_repeat", part0,
                                                  os.path.join(ring, "rd.sim.csv"), sc_file], threads=thr)
                variants.setdefault(tuple(sorted(lines)), []).append(thr)
    init_s = set(init)
    news = {tuple(l for l in v if l not in init_s) for v in variants}
    assert len(news) == 1, "the new lines differ between runs"
    p = os.path.join(HERE, "ring_diff.json")
    g = json.load(open(p))
    g["sc_lines"] = [l for l in open(sc_file).read().splitlines() if l]
    g["sc_variants"] = [{"lines": list(v), "threads": t} for v, t in sorted(variants.items(), key=lambda kv: len(kv[0]))]
    g["sc_new_lines"] = list(news.pop())
    json.dump(g, open(p, "w"), indent=1)
    print("SC variants:", [(len(v["lines"]), v["threads"]) for v in g["sc_variants"]], "new lines:", len(g["sc_new_lines"]))


if __name__ == "__main__":
    main()
