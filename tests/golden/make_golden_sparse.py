#!/usr/bin/env python
"""Golden fixture for the --diff_init store on networks where some LPs never execute an event, from the UNMODIFIED
reference (build container only: needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden_sparse.py
The reference stores / fossil-collects only LPs its scheduler ever dequeued (runner::clear walks
scheduler::active_lp, runner.hpp:573-583): an LP no event reached has NO record in the state store, not even its
initial (lp, -1, -1) version. Found by tools/emu/fuzz.py (mode refstore); the ring and grid goldens do not show it
because every LP is visited there. Two small cases: 4 junctions / 1 trip (junction 1 is never visited) and
40 junctions / 25 trips. Every trip ends well before finish_time, so nothing depends on the reference's thread timing.
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
from make_golden import run_ref  # noqa: E402
from make_golden_diff import store_keys  # noqa: E402
from fixtures import sparse_network as network  # noqa: E402
from scalesim_b200 import scenario  # noqa: E402


def main():
    out = {"source": "oracle/_ref/TrafficSim --diff_init on tests/fixtures.sparse_network(n_nodes, n_trips, seed)", "cases": []}
    for n_nodes, n_trips, seed in ((4, 1, 3), (40, 25, 8)):
        sc = network(n_nodes, n_trips, seed)
        with tempfile.TemporaryDirectory() as td:
            rd, trip, part = (os.path.join(td, n) for n in ("rd.csv", "trip.csv", "part"))
            scenario.write_traffic_csv(sc, rd, trip, part)
            st = os.path.join(td, "st")
            os.makedirs(st)
            runs = []
            for thr in (2, 3, 4):
                lines, _ = run_ref("TrafficSim", ["--sim_id=7", f"--store_dir={st}/{thr}", "--diff_init", part, rd, trip], threads=thr)
                runs.append((lines, {db: store_keys(os.path.join(st, str(thr), "7", db + "0"))[0] for db in ("event", "cancel", "state")}))
            assert all(r == runs[0] for r in runs), "the store depends on the thread count"
        lines, keys = runs[0]
        visited = sorted(set(k[0] for k in keys["event"]))
        assert len(visited) < n_nodes, "every LP is visited: the case shows nothing"
        assert sorted(k[0] for k in keys["state"] if k[1] == -1) == visited
        out["cases"].append({"n_nodes": n_nodes, "n_trips": n_trips, "seed": seed, "init_lines": lines, "store": keys,
                             "visited_lps": visited})
        print(n_nodes, n_trips, seed, "lines", len(lines), {k: len(v) for k, v in keys.items()}, "visited", len(visited), "of", n_nodes)
    json.dump(out, open(os.path.join(HERE, "sparse_diff.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
