"""Host-side mirror of ScaleSim's TrafficSim application (reference src/trafficsim/traffic_sim.hpp):
the file readers, a synthetic grid generator, and the packing into the engine's records.
The handler runs on the GPU (``TrafficModel`` in csrc/ssb_models.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .capi import EVENT_DTYPE, FLAG_HISTORY, MODEL_TRAFFIC, Config, Engine

FINISH_TIME = 10800      # traffic_sim.hpp:240-247
INLINE_ROADS = 2         # TrafficModel::kInlineRoads: junctions with more out-roads keep them in the road table
STATE_WORDS = 2 + 3 * INLINE_ROADS


@dataclass
class Scenario:
    """Road network + trips in flat arrays."""
    n_lp: int
    state_ids: np.ndarray     # LP id of each state record (quirk Q10: field 2 of the LAST road on the line)
    road_start: np.ndarray    # CSR over state records
    road_dst: np.ndarray
    road_speed: np.ndarray
    road_len: np.ndarray
    trip_id: np.ndarray
    trip_dep: np.ndarray
    route_start: np.ndarray   # CSR over trips
    route_nodes: np.ndarray
    partition: np.ndarray | None = None


def _atol(s: str) -> int:
    """C atol: optional whitespace, sign, digits; anything else ends the number; no digits = 0."""
    s = s.lstrip(" \t\n\v\f\r")
    i, sign = 0, 1
    if s[:1] in "+-" and s[:1]:
        sign = -1 if s[0] == "-" else 1
        i = 1
    j = i
    while j < len(s) and "0" <= s[j] <= "9":      # ASCII digits only (str.isdigit also accepts other scripts' digits)
        j += 1
    return sign * int(s[i:j]) if j > i else 0


def _lines(path: str) -> list:
    """The lines std::getline yields: split at '\n' only (a '\r' stays in the line, where atol ignores it), the last
    line counts with or without a final '\n'."""
    with open(path, newline="", encoding="latin-1") as f:
        text = f.read()
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def read_partition(path: str) -> np.ndarray:
    """traffic_reader::graph_read (traffic_sim.hpp:345-362): one partition number per line, line index = LP id."""
    return np.array([_atol(line) for line in _lines(path)], dtype=np.int64)


def read_roads(path: str):
    """traffic_reader::road_read (traffic_sim.hpp:364-408), all partitions. One line per LP; roads split on ';',
    fields on ','; state id = field 2 of the last road; length 0 -> 1; lanes ignored by the handler."""
    state_ids, start, dst, speed, length = [], [0], [], [], []
    for line in _lines(path):
        sid = -1
        for road in line.split(";"):
            fld = road.split(",")
            if len(fld) < 7:
                continue
            sid = _atol(fld[2])
            dst.append(_atol(fld[3]))
            speed.append(_atol(fld[4]))
            ln = _atol(fld[5])
            length.append(ln if ln != 0 else 1)
        state_ids.append(sid)
        start.append(len(dst))
    return (np.array(state_ids, dtype=np.int64), np.array(start, dtype=np.int64), np.array(dst, dtype=np.int64),
            np.array(speed, dtype=np.int64), np.array(length, dtype=np.int64))


def read_trips(path: str):
    """traffic_reader::trip_read (traffic_sim.hpp:410-451): vehicle id = field 1, arrival = departure = field 3,
    route = fields 4.., source = -1 (field 2 ignored)."""
    tid, dep, start, nodes = [], [], [0], []
    for line in _lines(path):
        fld = line.split(",")
        if len(fld) < 5:
            continue
        tid.append(_atol(fld[1]))
        dep.append(_atol(fld[3]))
        nodes.extend(_atol(x) for x in fld[4:])
        start.append(len(nodes))
    return (np.array(tid, dtype=np.int64), np.array(dep, dtype=np.int64), np.array(start, dtype=np.int64),
            np.array(nodes, dtype=np.int64))


@dataclass
class WhatIf:
    """What-if queries of a --diff_repeat run: added trips (AE) and deleted events (DE / RE)."""
    trip_id: np.ndarray
    trip_dep: np.ndarray
    route_start: np.ndarray
    route_nodes: np.ndarray
    deletes: np.ndarray          # rows (lp, time, event id)
    state_changes: list = None   # SC queries: (lp, time, road dst[], speed[], length[])

    def __iter__(self):          # unpacks like the (trip_id, trip_dep, route_start, route_nodes) tuple of the AE part
        return iter((self.trip_id, self.trip_dep, self.route_start, self.route_nodes))


def read_what_if(path: str) -> WhatIf:
    """traffic_reader::what_if_read (traffic_sim.hpp:453-599).
    ``AE,{event id},0,{departure},{1st node},{2nd node},...`` add event: query LP = first node, time = departure (:522-558)
    ``DE,{lp id},{time},{event id}``                           delete event (:559-580)
    ``RE,{event id},{source},{send time},{destination},{receive time}``   delete the event of that result line (:581-596)
    ``SC,{lp id},{time};{road};{road};...``                    state change: the LP's roads from ``time`` on (:473-521);
                                                               road = ``R,{id},{lp},{dst},{speed},{length},{lanes}``"""
    tid, dep, start, nodes, dels, scs = [], [], [0], [], [], []
    for line in _lines(path):
        fld = line.split(",")
        if line[:2] == "AE":
            tid.append(_atol(fld[1]))
            dep.append(_atol(fld[3]))
            nodes.extend(_atol(x) for x in fld[4:])
            start.append(len(nodes))
        elif line[:2] == "DE":
            dels.append((_atol(fld[1]), _atol(fld[2]), _atol(fld[3])))
        elif line[:2] == "RE":
            dels.append((_atol(fld[4]), _atol(fld[5]), _atol(fld[1])))
        elif line[:2] == "SC":
            parts = line.split(";")
            head = parts[0].split(",")
            roads = [r.split(",") for r in parts[1:]]
            ln = [_atol(r[5]) for r in roads]
            scs.append((_atol(head[1]), _atol(head[2]), [_atol(r[3]) for r in roads], [_atol(r[4]) for r in roads],
                        [x if x != 0 else 1 for x in ln]))
    return WhatIf(np.array(tid, dtype=np.int64), np.array(dep, dtype=np.int64), np.array(start, dtype=np.int64),
                  np.array(nodes, dtype=np.int64), np.array(dels, dtype=np.int64).reshape(-1, 3), scs)


def read_scenario(road_csv: str, trip_csv: str, partition_file: str | None = None) -> Scenario:
    sid, rs, rd, sp, ln = read_roads(road_csv)
    tid, dep, ts, nodes = read_trips(trip_csv)
    part = read_partition(partition_file) if partition_file else None
    n_lp = int(max(sid.max(initial=-1), rd.max(initial=-1), nodes.max(initial=-1))) + 1
    if part is not None:
        n_lp = max(n_lp, part.size)
    return Scenario(n_lp, sid, rs, rd, sp, ln, tid, dep, ts, nodes, part)


# ---------------------------------------------------------------------------------------
# synthetic grid (BASELINE.json config 3; the reference ships no generator — SURVEY.md §8d freezes this one)
# ---------------------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_grid(width: int = 1024, height: int = 1024, n_trips: int = 10_000_000, seed: int = 1,
                   min_hops: int = 8, max_hops: int = 64, max_departure: int = 7200) -> Scenario:
    """Torus grid: node = r*W + c, two out-roads per node (right, down, with wrap), road id = 2*node + {0,1};
    speed in {10..60} km/h and length in {100..1000} m from splitmix64(road_id ^ seed); trip t: departure =
    h % 7200, start = h' % (W*H), hop count in [8, 64], each hop right/down from hash bits."""
    n = width * height
    node = np.arange(n, dtype=np.int64)
    r, c = node // width, node % width
    right = r * width + (c + 1) % width
    down = ((r + 1) % height) * width + c
    road_dst = np.stack([right, down], axis=1).reshape(-1)
    rid = np.arange(2 * n, dtype=np.uint64)
    h = _splitmix64(rid ^ np.uint64(seed))
    road_speed = (10 + (h % np.uint64(51))).astype(np.int64)
    road_len = (100 + ((h >> np.uint64(20)) % np.uint64(901))).astype(np.int64)
    road_start = np.arange(0, 2 * n + 1, 2, dtype=np.int64)

    t = np.arange(n_trips, dtype=np.uint64)
    h1 = _splitmix64(t ^ (np.uint64(seed) << np.uint64(32)))
    h2 = _splitmix64(h1)
    h3 = _splitmix64(h2)
    dep = (h1 % np.uint64(max_departure)).astype(np.int64)
    start = (h2 % np.uint64(n)).astype(np.int64)
    hops = (min_hops + (h3 % np.uint64(max_hops - min_hops + 1))).astype(np.int64)   # number of moves
    lens = hops + 1
    route_start = np.zeros(n_trips + 1, dtype=np.int64)
    np.cumsum(lens, out=route_start[1:])
    nodes = np.empty(int(route_start[-1]), dtype=np.int64)
    # direction bits: bit k of splitmix64(h3 + k // 64)
    cur_r, cur_c = start // width, start % width
    alive = np.ones(n_trips, dtype=bool)
    bits = _splitmix64(h3)
    for k in range(int(max_hops) + 1):
        idx = np.nonzero(alive)[0]
        if idx.size == 0:
            break
        nodes[route_start[idx] + k] = cur_r[idx] * width + cur_c[idx]
        if k % 64 == 0 and k > 0:
            bits = _splitmix64(bits)
        go_down = ((bits >> np.uint64(k % 64)) & np.uint64(1)).astype(bool)
        mv = alive & (k < hops)
        cur_r = np.where(mv & go_down, (cur_r + 1) % height, cur_r)
        cur_c = np.where(mv & ~go_down, (cur_c + 1) % width, cur_c)
        alive = alive & (k + 1 < lens)
    part = None
    return Scenario(n, node.copy(), road_start, road_dst, road_speed, road_len, t.astype(np.int64), dep,
                    route_start, nodes, part)


def stripe_partition(width: int, height: int, world: int) -> np.ndarray:
    """Row-stripe partition of the grid (METIS is not available): rows [k*H/world, (k+1)*H/world) -> k."""
    rows = np.arange(height, dtype=np.int64) * world // height
    return np.repeat(rows, width)


# ---------------------------------------------------------------------------------------
# packing
# ---------------------------------------------------------------------------------------
def pack_states(sc: Scenario):
    """LP state words [id, n_roads, ...]: up to INLINE_ROADS roads inline as (dst, speed, length) triples; a junction
    with more keeps the index of its first road in the road table (model data slot 1) in word 2.
    Returns (state LP ids, words, road table as int64 triples)."""
    n = sc.state_ids.size
    deg = np.diff(sc.road_start)
    words = np.zeros((n, STATE_WORDS), dtype=np.uint32)
    words[:, 0] = sc.state_ids.astype(np.uint32)
    words[:, 1] = deg.astype(np.uint32)
    small = deg <= INLINE_ROADS
    for k in range(INLINE_ROADS):
        has = small & (deg > k)
        j = sc.road_start[:-1][has] + k
        words[has, 2 + 3 * k] = sc.road_dst[j].astype(np.uint32)
        words[has, 3 + 3 * k] = sc.road_speed[j].astype(np.uint32)
        words[has, 4 + 3 * k] = sc.road_len[j].astype(np.uint32)
    big = np.nonzero(~small)[0]
    table = np.zeros((0, 3), dtype=np.int64)
    if big.size:
        first = np.zeros(big.size, dtype=np.int64)
        np.cumsum(deg[big][:-1], out=first[1:])
        words[big, 2] = first.astype(np.uint32)
        rows = np.concatenate([np.arange(sc.road_start[i], sc.road_start[i + 1]) for i in big])
        table = np.stack([sc.road_dst[rows], sc.road_speed[rows], sc.road_len[rows]], axis=1).astype(np.int64)
    return sc.state_ids.astype(np.int64), words, table


def pack_events(sc: Scenario) -> np.ndarray:
    """One initial event per trip: source -1, destination = first route node, p0 = offset of that node in the
    route pool, p1 = route length."""
    n = sc.trip_id.size
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    ev["id"] = sc.trip_id
    ev["src"] = -1
    ev["dst"] = sc.route_nodes[sc.route_start[:-1]]
    ev["recv_time"] = sc.trip_dep
    ev["send_time"] = sc.trip_dep
    ev["p0"] = sc.route_start[:-1]
    ev["p1"] = np.diff(sc.route_start)
    return ev


def make_engine(sc: Scenario, *, bucket_ticks: int = 8, window_buckets: int = 1, keep_records: bool = False,
                rank: int = 0, world: int = 1, device: int = 0, max_events: int | None = None,
                max_batch: int | None = None, max_committed: int | None = None, finish_time: int = FINISH_TIME,
                rounds_per_sync: int = 8, load: bool = True, history: bool = False, extra_routes=None,
                extra_roads=None) -> Engine:
    """Build a TrafficSim engine the way runner<traffic_sim>::init does (runner.hpp:79-176)."""
    n_trips = sc.trip_id.size
    n_nodes = sc.route_nodes.size
    if max_events is None:
        max_events = int(n_trips * 1.5) + (1 << 16) + 4096 * 8
    if max_batch is None:
        max_batch = max(1 << 14, n_trips // 2 + (1 << 14))
    if max_committed is None:
        max_committed = n_nodes + 1024 if (keep_records or history) else 0
    cfg = Config(model=MODEL_TRAFFIC, n_lp=sc.n_lp, finish_time=finish_time, bucket_ticks=bucket_ticks,
                 window_buckets=window_buckets, max_events=max_events, max_batch=max_batch,
                 max_committed=max_committed, rank=rank, world=world, device=device, rounds_per_sync=rounds_per_sync,
                 flags=FLAG_HISTORY if history else 0)
    eng = Engine(cfg)
    part = sc.partition
    if part is not None and part.size < sc.n_lp:
        part = np.concatenate([part, np.zeros(sc.n_lp - part.size, dtype=np.int64)])
    eng.set_partition(part % world if part is not None else None)
    pool = sc.route_nodes.astype(np.int64)
    if extra_routes is not None and len(extra_routes):
        pool = np.concatenate([pool, np.asarray(extra_routes, dtype=np.int64)])      # what-if trips: routes appended
    eng.set_model_data(0, pool)
    ids, words, road_table = pack_states(sc)
    if extra_roads is not None and len(extra_roads):
        road_table = np.concatenate([road_table, np.asarray(extra_roads, dtype=np.int64).reshape(-1, 3)])
    if road_table.size:
        eng.set_model_data(1, road_table.reshape(-1))
    owner = (part[ids] % world) if part is not None else (ids % world)
    mine = owner == rank
    eng.load_states(ids[mine], words[mine])
    if load:
        ev = pack_events(sc)
        ev_owner = (part[ev["dst"]] % world) if part is not None else (ev["dst"] % world)
        eng.load_events(ev[ev_owner == rank])
    return eng


# ---------------------------------------------------------------------------------------
# exact-differential simulation (reference --diff_init / --diff_repeat)
# ---------------------------------------------------------------------------------------
def run_diff_init(sc: Scenario, **kw):
    """--diff_init: a normal run that keeps the history. Returns (engine stats, committed records, HistoryStore)."""
    from .store import HistoryStore
    rank, world = kw.get("rank", 0), kw.get("world", 1)
    eng = make_engine(sc, history=True, **kw)
    st = eng.run()
    rec, sent = eng.drain_history()
    eng.close()
    part = sc.partition
    owner = (part[sc.state_ids] % world) if part is not None else (sc.state_ids % world)
    return st, rec, HistoryStore(rec, sent, sc.state_ids[owner == rank], sc.route_nodes)


def run_diff_init_streaming(sc: Scenario, store_dir: str, sim_id, **kw):
    """--diff_init with the history leaving the device WHILE the window loop runs (ssb_run_stream_history): chunks go
    from the engine's pinned staging buffers to a writer thread that appends them to the store file. Returns
    (engine stats, path of the store)."""
    from .store import HistoryStore, HistoryWriter
    rank, world = kw.get("rank", 0), kw.get("world", 1)
    eng = make_engine(sc, history=True, **kw)
    part = sc.partition
    owner = (part[sc.state_ids] % world) if part is not None else (sc.state_ids % world)
    w = HistoryWriter(HistoryStore.path(store_dir, sim_id, rank), sc.state_ids[owner == rank], sc.route_nodes)
    st = eng.run_stream_history(lambda rec, sent: w.append(rec, sent))
    eng.close()
    return st, w.close()


def make_repeat_engine(sc: Scenario, store, what_if, **kw) -> Engine:
    """--diff_repeat (runner::init_repeat, runner.hpp:178-348): the recorded history goes back into the engine as
    processed events, then the AE queries are injected. ``sc`` provides the road network and the ORIGINAL route pool
    (history events index into it); ``what_if`` = (trip_id, trip_dep, route_start, route_nodes) from read_what_if.
    ``engine.run()`` then re-executes exactly what the queries invalidate and commits only that."""
    rank, world = kw.get("rank", 0), kw.get("world", 1)
    tid, dep, rstart, rnodes = what_if
    deletes = getattr(what_if, "deletes", np.zeros((0, 3), dtype=np.int64))
    state_changes = getattr(what_if, "state_changes", None) or []
    n_hist = store.rec.size
    # re-executed history is re-sent: until a bucket is fossil-collected it holds the recorded copy AND the new one
    kw.setdefault("max_events", int((n_hist + rnodes.size) * 3) + (1 << 16))
    kw.setdefault("max_batch", max(1 << 14, (n_hist + tid.size) // 2 + (1 << 14)))
    kw.setdefault("max_committed", n_hist + int(rnodes.size) + 1024)
    # state-change queries: the new road sets are packed like states (junctions with more than INLINE_ROADS roads go
    # to the road table, behind the scenario's own entries)
    sc_words = []
    if state_changes:
        _, base_words, base_table = pack_states(sc)
        extra_rows = []
        for lp, t, dst, speed, length in state_changes:
            w = np.zeros(STATE_WORDS, dtype=np.uint32)
            w[0], w[1] = lp, len(dst)
            if len(dst) <= INLINE_ROADS:
                for k in range(len(dst)):
                    w[2 + 3 * k: 5 + 3 * k] = (dst[k], speed[k], length[k])
            else:
                w[2] = base_table.shape[0] + len(extra_rows)
                extra_rows.extend(zip(dst, speed, length))
            sc_words.append(w)
        kw = dict(kw, extra_roads=np.array(extra_rows, dtype=np.int64).reshape(-1, 3))
    eng = make_engine(sc, load=False, extra_routes=rnodes, **kw)
    sent = store.sent
    if deletes.size:
        # delete event (runner.hpp:246-276): the recorded event loses its sent-log entry — eventq::delete_can,
        # queue.hpp:238-241: its output is NOT cancelled — and an anti-message for it goes into the LP's inbox
        # (eventq::delete_ev, :227-236), which erases it and rolls the LP back to its key
        sent = sent.copy()
        keys = (deletes[:, 1].astype(np.uint64) << np.uint64(32)) | deletes[:, 2].astype(np.uint64)
        # one sort of the recorded keys, then a binary search per query (not a scan of the history per query)
        order = np.argsort(store.rec["key"], kind="stable")
        sorted_keys = store.rec["key"][order]
        lo = np.searchsorted(sorted_keys, keys, side="left")
        hi = np.searchsorted(sorted_keys, keys, side="right")
        for lp, a, b in zip(deletes[:, 0], lo, hi):
            hit = order[a:b]
            sent["dst"][hit[store.rec["dst"][hit] == lp]] = 0xFFFFFFFF
    eng.load_history(store.rec, sent)
    part = sc.partition
    for (lp, t, *_), w in zip(state_changes, sc_words):
        if ((part[lp] % world) if part is not None else (lp % world)) == rank:
            eng.set_state_from(lp, t, w)
    ev = np.zeros(tid.size, dtype=EVENT_DTYPE)
    ev["id"] = tid
    ev["src"] = -1
    ev["dst"] = rnodes[rstart[:-1]]
    ev["recv_time"] = dep
    ev["send_time"] = dep
    ev["p0"] = sc.route_nodes.size + rstart[:-1]
    ev["p1"] = np.diff(rstart)
    if deletes.size:
        anti = np.zeros(deletes.shape[0], dtype=EVENT_DTYPE)
        anti["id"] = deletes[:, 2]
        anti["src"] = -1
        anti["dst"] = deletes[:, 0]
        anti["recv_time"] = anti["send_time"] = deletes[:, 1]
        anti["flags"] = 1
        ev = np.concatenate([ev, anti])
    owner = (part[ev["dst"]] % world) if part is not None else (ev["dst"] % world)
    eng.load_events(ev[owner == rank])
    return eng
