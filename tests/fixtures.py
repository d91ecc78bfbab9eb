"""Shared helpers for the tests: golden fixtures and scenario builders."""
import gzip
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    if "digest" in g:
        g["digest"] = tuple(int(x) for x in g["digest"])
    return g


def golden_lines(name):
    with gzip.open(os.path.join(GOLDEN, name + ".lines.gz"), "rt") as f:
        return f.read().splitlines()


def sha256_lines(lines):
    return hashlib.sha256("".join(l + "\n" for l in sorted(lines)).encode()).hexdigest()


def ring_scenario():
    from scalesim_b200 import traffic
    g = golden("ring")
    a = lambda k: np.array(g[k], dtype=np.int64)
    return traffic.Scenario(g["n_lp"], a("state_ids"), a("road_start"), a("road_dst"), a("road_speed"), a("road_len"),
                            a("trip_id"), a("trip_dep"), a("route_start"), a("route_nodes"), a("partition4")), g


def write_ring_files(tmpdir):
    """Re-create the ring inputs in the reference's CSV formats (traffic/README.md) from the array fixture."""
    sc, g = ring_scenario()
    rd = os.path.join(tmpdir, "rd.csv")
    trip = os.path.join(tmpdir, "trip.csv")
    part = os.path.join(tmpdir, "graph.part.4")
    with open(rd, "w") as f:
        rid = 0
        for i, sid in enumerate(sc.state_ids):
            roads = []
            for j in range(sc.road_start[i], sc.road_start[i + 1]):
                roads.append(f"R,{rid},{sid},{sc.road_dst[j]},{sc.road_speed[j]},{sc.road_len[j]},1")
                rid += 1
            f.write(";".join(roads) + "\n")
    with open(trip, "w") as f:
        for t in range(sc.trip_id.size):
            route = ",".join(str(x) for x in sc.route_nodes[sc.route_start[t]:sc.route_start[t + 1]])
            f.write(f"TP,{sc.trip_id[t]},0,{sc.trip_dep[t]},{route}\n")
    with open(part, "w") as f:
        f.write("".join(f"{p}\n" for p in sc.partition))
    return rd, trip, part


def sparse_network(n_nodes, n_trips, seed):
    """tests/golden/sparse_diff.json: random network in which some junctions are never visited; speeds 20..60 (no zero
    speed: no hop reaches finish_time), short random walks"""
    from scalesim_b200 import traffic
    rng = np.random.default_rng(seed)
    deg = rng.integers(1, 3, n_nodes)
    road_start = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(deg, out=road_start[1:])
    road_dst = rng.integers(0, n_nodes, road_start[-1])
    road_speed = rng.integers(20, 61, road_start[-1])
    road_len = rng.integers(100, 1200, road_start[-1])
    dep, start, nodes = [], [0], []
    for _ in range(n_trips):
        cur = int(rng.integers(0, n_nodes))
        route = [cur]
        for _ in range(int(rng.integers(1, 6))):
            cur = int(road_dst[road_start[cur] + rng.integers(0, deg[cur])])
            route.append(cur)
        dep.append(int(rng.integers(0, 3000)))
        nodes.extend(route)
        start.append(len(nodes))
    a = lambda x: np.array(x, dtype=np.int64)
    return traffic.Scenario(n_nodes, np.arange(n_nodes, dtype=np.int64), road_start, road_dst, road_speed, road_len,
                            np.arange(n_trips, dtype=np.int64), a(dep), a(start), a(nodes), np.zeros(n_nodes, dtype=np.int64))
