"""Pins the oracle on the reference beyond the committed goldens: on seeded random inputs the UNMODIFIED reference
binaries (oracle/_ref/PHOLD, oracle/_ref/TrafficSim — built from /root/reference where that tree exists, skipped
elsewhere) and the oracle's restatement print the same lines. (tools/emu/fuzz.py's `ref` mode does the same in bulk,
with the drop-in binary on the emulated kernels as the third party.)"""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from scalesim_b200 import scenario, traffic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _ref_lines(exe, args, threads):
    path = os.path.join(REF, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the reference tree)")
    for _ in range(5):      # intermittent start-up crash of the reference: retry
        p = subprocess.run([path] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                           env=dict(os.environ, SCALESIM_NUM_THR=str(threads)))
        if p.returncode == 0:
            return sorted(l for l in p.stdout.splitlines() if l.startswith("RE,"))
    pytest.fail(f"{exe} kept failing: {p.stderr[-500:]}")


@pytest.mark.parametrize("n_lp,n_init,ratio,lam,threads", [
    (1, 300, 1.0, 3.0, 1), (8, 40, 0.5, 0.5, 2), (50, 600, 0.0, 1.0, 4), (120, 900, 0.3, 2.0, 2), (200, 1, 0.1, 1.0, 1),
])
def test_phold_oracle_equals_reference_binary(n_lp, n_init, ratio, lam, threads):
    ref = _ref_lines("PHOLD", [n_lp, n_init, ratio, lam, 0, 50], threads)
    s = O.Sim.phold(n_lp, n_init, ratio, lam)
    s.run_closure(2)
    assert sorted(O.records_to_lines(s.committed())) == ref
    s2 = O.Sim.phold(n_lp, n_init, ratio, lam)
    s2.run_timewarp(3, 50, 2)                       # the restated Time Warp loop, not only the closure
    assert sorted(O.records_to_lines(s2.committed())) == ref


def _network(n_nodes, max_degree, n_trips, seed):
    """random directed road network (every junction has 1..max_degree out-roads, zero speeds occur) with random walks
    as trips, some ending in a node no road leads to"""
    rng = np.random.default_rng(seed)
    deg = rng.integers(1, max_degree + 1, n_nodes)
    road_start = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(deg, out=road_start[1:])
    road_dst = rng.integers(0, n_nodes, road_start[-1])
    road_speed = rng.integers(0, 61, road_start[-1])
    road_len = rng.integers(1, 1200, road_start[-1])
    dep, start, nodes = [], [0], []
    for _ in range(n_trips):
        cur = int(rng.integers(0, n_nodes))
        route = [cur]
        for _ in range(int(rng.integers(1, 12))):
            cur = int(road_dst[road_start[cur] + rng.integers(0, deg[cur])])
            route.append(cur)
        if rng.random() < 0.1:
            route.append(int(rng.integers(0, n_nodes)))
        dep.append(int(rng.integers(0, 3000)))
        nodes.extend(route)
        start.append(len(nodes))
    a = lambda x: np.array(x, dtype=np.int64)
    return traffic.Scenario(n_nodes, np.arange(n_nodes, dtype=np.int64), road_start, road_dst, road_speed, road_len,
                            np.arange(n_trips, dtype=np.int64), a(dep), a(start), a(nodes), np.zeros(n_nodes, dtype=np.int64))


@pytest.mark.parametrize("n_nodes,deg,n_trips,seed,threads", [(4, 2, 30, 11, 2), (30, 5, 400, 12, 4), (150, 1, 60, 13, 2), (40, 9, 200, 14, 3)])
def test_traffic_oracle_equals_reference_binary(tmp_path, n_nodes, deg, n_trips, seed, threads):
    sc = _network(n_nodes, deg, n_trips, seed)
    rd, trip, part = (str(tmp_path / n) for n in ("rd.csv", "trip.csv", "part"))
    scenario.write_traffic_csv(sc, rd, trip, part)
    ref = _ref_lines("TrafficSim", [part, rd, trip], threads)
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(2)
    assert sorted(O.records_to_lines(s.committed())) == ref
    s2 = O.Sim.traffic_files(rd, trip)              # the oracle's own restatement of the file readers
    s2.run_timewarp(3, 50, 2)
    assert sorted(O.records_to_lines(s2.committed())) == ref
