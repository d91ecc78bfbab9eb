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
