"""Developer tool: randomized differential runs of the kernel LOGIC (host emulation, tools/emu/libssb_emu.so) against
the CPU oracle — PHOLD (plain and state-mutating) and TrafficSim on random road networks, with random calendar /
window / capacity settings, stream path and graph path. Finds logic errors where there is no GPU; parity claims are
made on the B200 only.

    tools/emu/build.sh && python tools/emu/fuzz.py [seconds] [seed]
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_plugin  # noqa: F401,E402  (points capi at the emulation)
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
import scalesim_b200 as ssb  # noqa: E402
from scalesim_b200 import phold, traffic  # noqa: E402
from test_gpu_traffic import _random_network  # noqa: E402

_tables = {}


def tables(ratio, lam, seed):
    k = (ratio, lam, seed)
    if k not in _tables:
        if len(_tables) > 6:
            _tables.clear()
        _tables[k] = phold.build_tables(ratio, lam, seed=seed)
    return _tables[k]


def one_phold(rng):
    n_lp = int(rng.choice([1, 2, 7, 64, 300, 1500]))
    n_init = int(rng.choice([1, 5, 200, 3000, 12000]))
    ratio = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
    lam = float(rng.choice([0.5, 1.0, 4.0, 30.0]))
    seed = int(rng.choice([1, 2, 3]))
    stateful = bool(rng.random() < 0.4)
    bucket = int(rng.choice([1, 50, 1000, 2500, 10000, 25000, 200000]))
    window = int(rng.choice([1, 2, 3, 8, 64]))
    if bucket * window < 50:
        window = 64
    if (phold.FINISH_TIME // bucket) > 20000:       # bound the number of emulated rounds
        bucket = 50
    cfg = dict(n_lp=n_lp, n_init=n_init, ratio=ratio, lam=lam, seed=seed, stateful=stateful, bucket=bucket, window=window,
               rps=int(rng.choice([1, 4, 8])), graph=bool(rng.random() < 0.5))
    if os.environ.get("FUZZ_VERBOSE"):
        print("phold", cfg, flush=True)
    s = O.Sim.phold(n_lp, n_init, ratio, lam, seed=seed, stateful=stateful)
    if stateful:
        s.run_sequential()
    else:
        s.run_closure(2)
    want = s.digest()
    if not cfg["graph"]:
        os.environ["SSB_NO_GRAPH"] = "1"
    try:
        eng = phold.make_engine(n_lp, n_init, ratio, lam, tables=tables(ratio, lam, seed), bucket_ticks=bucket, window_buckets=window,
                                stateful=stateful, rounds_per_sync=cfg["rps"], max_events=int(n_init * 3) + (1 << 16),
                                max_batch=n_init + (1 << 14), max_committed=0)
    finally:
        os.environ.pop("SSB_NO_GRAPH", None)
    st = eng.run()
    got = eng.digest()
    ok = got == want
    if ok and stateful:
        ids, words = s.final_states(n_lp)
        ok = bool(np.array_equal(eng.read_states(ids), words))
    if ok and rng.random() < 0.3:
        # the same engine again after ssb_reset (what bench.py does between steps): same result, same statistics
        eng.reset()
        eng.load_events(phold.init_events(n_lp, n_init, tables(ratio, lam, seed)[0]))
        st_b = eng.run()
        ok = eng.digest() == want and (st_b.committed, st_b.rollbacks, st_b.antis_sent) == (st.committed, st.rollbacks, st.antis_sent)
        cfg["reset"] = True
    eng.close()
    if not ok and stateful and st.duplicates > 0:
        # Two DIFFERENT events with one key (receive time, id): the state-mutating test model makes an event's output
        # depend on the state, so a stale trajectory and the re-executed one can collide. R1 keeps one event per key and
        # anti-messages match by key (queue.hpp:87-93) — the reference cannot tell such events apart either, so a
        # Time Warp run may then differ from the sequential one. Not an engine error; counted separately.
        return None, cfg, dict(collision=True)
    return ok, cfg, dict(committed=st.committed, rollbacks=st.rollbacks, antis=st.antis_sent, fix_turns=st.fix_turns, want=want[0])


def one_traffic(rng):
    n_nodes = int(rng.choice([3, 40, 300]))
    deg = int(rng.choice([1, 2, 3, 9]))
    n_trips = int(rng.choice([1, 50, 2000, 6000]))
    seed = int(rng.integers(0, 1 << 30))
    bucket = int(rng.choice([1, 5, 11, 64, 400, 4000]))
    window = int(rng.choice([1, 2, 4, 32]))
    cfg = dict(n_nodes=n_nodes, deg=deg, n_trips=n_trips, seed=seed, bucket=bucket, window=window, graph=bool(rng.random() < 0.5))
    if os.environ.get("FUZZ_VERBOSE"):
        print("traffic", cfg, flush=True)
    sc = _random_network(n_nodes, deg, n_trips, seed)
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(2)
    want = s.digest()
    if not cfg["graph"]:
        os.environ["SSB_NO_GRAPH"] = "1"
    try:
        eng = traffic.make_engine(sc, bucket_ticks=bucket, window_buckets=window)
    finally:
        os.environ.pop("SSB_NO_GRAPH", None)
    st = eng.run()
    ok = eng.digest() == want
    eng.close()
    return ok, cfg, dict(committed=st.committed, rollbacks=st.rollbacks, antis=st.antis_sent, fix_turns=st.fix_turns, want=want[0])


def one_diff(rng):
    """--diff_init, then --diff_repeat with added trips (AE) and deleted events (DE) on a random network: the repeat
    output must be the closure of the queries over the init output (tests/oracle_lib.repeat_closure), and — for AE only —
    init + new == one full run over all trips (vehicles do not interact)."""
    n_nodes = int(rng.choice([12, 100, 400]))
    deg = int(rng.choice([1, 2, 4]))
    n_trips = int(rng.choice([20, 500, 3000]))
    n_add = int(rng.choice([1, 5, 60]))
    seed = int(rng.integers(0, 1 << 30))
    bucket = int(rng.choice([5, 11, 64, 400]))
    window = int(rng.choice([1, 2, 8]))
    rbucket = int(rng.choice([5, 11, 64, 400]))
    rwindow = int(rng.choice([1, 2, 8]))
    cfg = dict(kind="diff", n_nodes=n_nodes, deg=deg, n_trips=n_trips, n_add=n_add, seed=seed, bucket=bucket, window=window,
               rbucket=rbucket, rwindow=rwindow)
    if os.environ.get("FUZZ_VERBOSE"):
        print("diff", cfg, flush=True)
    both = _random_network(n_nodes, deg, n_trips + n_add, seed)
    cut = int(both.route_start[n_trips])
    sc = traffic.Scenario(both.n_lp, both.state_ids, both.road_start, both.road_dst, both.road_speed, both.road_len,
                          both.trip_id[:n_trips], both.trip_dep[:n_trips], both.route_start[:n_trips + 1], both.route_nodes[:cut])
    add = (both.trip_id[n_trips:], both.trip_dep[n_trips:], both.route_start[n_trips:] - cut, both.route_nodes[cut:])
    big = dict(max_events=(both.route_nodes.size + 64) * 12 + (1 << 16), max_batch=both.route_nodes.size * 4 + (1 << 14))
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=bucket, window_buckets=window, **big)
    init_lines = sorted(ssb.records_to_lines(ssb.raw_to_records(rec)))
    s = O.Sim.traffic_scenario(both)
    s.run_closure(2)
    full_lines = sorted(O.records_to_lines(s.committed()))
    s1 = O.Sim.traffic_scenario(sc)
    s1.run_closure(2)
    ok = init_lines == sorted(O.records_to_lines(s1.committed()))
    new_lines = sorted(set(full_lines) - set(init_lines))
    if rng.random() < 0.4 and len(init_lines) > 3:
        # delete-event queries (DE, runner.hpp:246-276) on top of the added trips: random recorded events disappear
        k = int(rng.choice([1, 3]))
        picks = rng.choice(len(init_lines), size=k, replace=False)
        dels = []
        for j in picks:
            i_, src_, send_, dst_, recv_ = O.parse_line(init_lines[int(j)])
            dels.append((dst_, recv_, i_))
        cfg["deletes"] = dels
        wi = traffic.WhatIf(add[0], add[1], add[2], add[3], np.array(dels, dtype=np.int64).reshape(-1, 3))
        eng = traffic.make_repeat_engine(sc, hs, wi, bucket_ticks=rbucket, window_buckets=rwindow, keep_records=True,
                                         max_committed=(both.route_nodes.size + 64) * 4, **big)
        st2 = eng.run()
        rep = sorted(ssb.records_to_lines(eng.drain_committed()))
        eng.close()
        ok = ok and rep == O.repeat_closure(init_lines, new_lines, dels)
        return ok, cfg, dict(init=len(init_lines), new=len(new_lines), repeat=len(rep), rollbacks=st2.rollbacks)
    eng = traffic.make_repeat_engine(sc, hs, add, bucket_ticks=rbucket, window_buckets=rwindow, keep_records=True,
                                     max_committed=(both.route_nodes.size + 64) * 4, **big)
    st2 = eng.run()
    rep = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    ok = ok and sorted(set(rep) - set(init_lines)) == new_lines and len(set(rep)) == len(rep) and set(rep) <= set(full_lines)
    ok = ok and rep == O.repeat_closure(init_lines, new_lines)
    return ok, cfg, dict(init=len(init_lines), new=len(new_lines), repeat=len(rep), rollbacks=st2.rollbacks)


def one_multi(rng):
    """Several partitions (tools/emu/multi.py: one emulated engine per host thread, mailboxes exchanged by address):
    PHOLD with lp % world, TrafficSim with a random partition table; the union of the partitions' committed multisets
    must be the oracle's."""
    import multi
    world = int(rng.choice([2, 3, 4, 8]))
    window = int(rng.choice([1, 2, 4, 16]))
    p2p = bool(rng.random() < 0.75)      # else the NCCL exchange mode (stream path; the emulation's in-process NCCL)
    graph = p2p and bool(rng.random() < 0.6)
    sync = int(rng.choice([1, 1, 2, 3, 8])) if graph else 1
    if rng.random() < 0.6:
        n_lp = int(rng.choice([world, 64, 300, 1500]))
        n_init = int(rng.choice([5, 200, 3000, 12000]))
        ratio = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
        lam = float(rng.choice([0.5, 1.0, 4.0, 30.0]))
        stateful = bool(rng.random() < 0.3)
        bucket = int(rng.choice([1000, 2500, 10000, 25000]))
        cfg = dict(kind="multi-phold", p2p=p2p, graph=graph, sync=sync, world=world, n_lp=n_lp, n_init=n_init, ratio=ratio, lam=lam, stateful=stateful, bucket=bucket, window=window)
        if os.environ.get("FUZZ_VERBOSE"):
            print("multi", cfg, flush=True)
        s = O.Sim.phold(n_lp, n_init, ratio, lam, stateful=stateful)
        if stateful:
            s.run_sequential()
        else:
            s.run_closure(2)
        want = s.digest()
        tb = tables(ratio, lam, 1)
        got, st, engines = multi.run_partitions(lambda r, w: phold.make_engine(
            n_lp, n_init, ratio, lam, tables=tb, rank=r, world=w, bucket_ticks=bucket, window_buckets=window, stateful=stateful,
            max_events=int(n_init * 3) + (1 << 16), max_batch=n_init + (1 << 14)), world, graph=graph, sync_interval=sync, keep_engines=True, p2p=p2p)
        if got == want and rng.random() < 0.3:
            # all partitions again after ssb_reset (bench.py between steps; the mailboxes are cleared too)
            for e in engines:
                e.reset()
            if stateful:
                for r, e in enumerate(engines):
                    e.load_states(*phold.init_states(n_lp, r, world, True))
            for r, e in enumerate(engines):
                e.load_events(phold.init_events(n_lp, n_init, tb[0], r, world))
            multi.run_all(engines)
            from scalesim_b200 import dist
            got = dist.combine_digests([e.digest() for e in engines])
            cfg["reset"] = True
        for e in engines:
            e.close()
        if got != want and stateful and sum(x.duplicates for x in st) > 0:
            return None, cfg, dict(collision=True)
    else:
        n_nodes = int(rng.choice([8, 40, 300]))
        deg = int(rng.choice([1, 2, 3, 9]))
        n_trips = int(rng.choice([50, 2000, 6000]))
        seed = int(rng.integers(0, 1 << 30))
        bucket = int(rng.choice([5, 11, 64, 400]))
        cfg = dict(kind="multi-traffic", p2p=p2p, graph=graph, sync=sync, world=world, n_nodes=n_nodes, deg=deg, n_trips=n_trips, seed=seed, bucket=bucket, window=window)
        if os.environ.get("FUZZ_VERBOSE"):
            print("multi", cfg, flush=True)
        sc = _random_network(n_nodes, deg, n_trips, seed)
        sc.partition = np.random.default_rng(seed).integers(0, 2 * world, n_nodes)      # reduced modulo world on load
        s = O.Sim.traffic_scenario(sc)
        s.run_closure(2)
        want = s.digest()
        got, st = multi.run_partitions(lambda r, w: traffic.make_engine(sc, rank=r, world=w, bucket_ticks=bucket, window_buckets=window,
                                                                        max_events=n_trips * 6 + (1 << 16), max_batch=n_trips * 2 + (1 << 14)),
                                     world, graph=graph, sync_interval=sync, p2p=p2p)
    return got == want, cfg, dict(committed=sum(x.committed for x in st), rollbacks=sum(x.rollbacks for x in st),
                                  antis=sum(x.antis_sent for x in st), remote=sum(x.remote_sent for x in st), want=want[0])


def one_multidiff(rng):
    """--diff_init and --diff_repeat with the LPs split over several partitions (one history store per partition): the
    union of the partitions' repeat outputs must be the closure of the added trips over the union of the init outputs."""
    import multi
    from scalesim_b200.store import HistoryStore
    world = int(rng.choice([2, 3, 4]))
    n_nodes = int(rng.choice([12, 100, 400]))
    deg = int(rng.choice([1, 2, 4]))
    n_trips = int(rng.choice([20, 500, 3000]))
    n_add = int(rng.choice([1, 5, 60]))
    seed = int(rng.integers(0, 1 << 30))
    bucket = int(rng.choice([5, 11, 64]))
    rbucket = int(rng.choice([5, 11, 64]))
    rwindow = int(rng.choice([1, 2, 8]))
    graph = bool(rng.random() < 0.5)
    cfg = dict(kind="multidiff", world=world, n_nodes=n_nodes, deg=deg, n_trips=n_trips, n_add=n_add, seed=seed, bucket=bucket,
               rbucket=rbucket, rwindow=rwindow, graph=graph)
    if os.environ.get("FUZZ_VERBOSE"):
        print("multidiff", cfg, flush=True)
    both = _random_network(n_nodes, deg, n_trips + n_add, seed)
    part = np.random.default_rng(seed).integers(0, world, n_nodes)
    cut = int(both.route_start[n_trips])
    sc = traffic.Scenario(both.n_lp, both.state_ids, both.road_start, both.road_dst, both.road_speed, both.road_len,
                          both.trip_id[:n_trips], both.trip_dep[:n_trips], both.route_start[:n_trips + 1], both.route_nodes[:cut], part)
    add = (both.trip_id[n_trips:], both.trip_dep[n_trips:], both.route_start[n_trips:] - cut, both.route_nodes[cut:])
    big = dict(max_events=(both.route_nodes.size + 64) * 12 + (1 << 16), max_batch=both.route_nodes.size * 4 + (1 << 14))
    _, _, engines = multi.run_partitions(lambda r, w: traffic.make_engine(sc, rank=r, world=w, bucket_ticks=bucket, history=True, **big),
                                         world, keep_engines=True, graph=graph)
    stores, init_lines = [], []
    for e in engines:
        rec, sent = e.drain_history()
        e.close()
        stores.append(HistoryStore(rec, sent, None, sc.route_nodes))
        init_lines += ssb.records_to_lines(ssb.raw_to_records(rec))
    init_lines.sort()
    s1 = O.Sim.traffic_scenario(sc)
    s1.run_closure(2)
    ok = init_lines == sorted(O.records_to_lines(s1.committed()))
    s = O.Sim.traffic_scenario(both)
    s.run_closure(2)
    new_lines = sorted(set(O.records_to_lines(s.committed())) - set(init_lines))
    _, st2, engines = multi.run_partitions(
        lambda r, w: traffic.make_repeat_engine(sc, stores[r], add, rank=r, world=w, bucket_ticks=rbucket, window_buckets=rwindow,
                                                keep_records=True, max_committed=(both.route_nodes.size + 64) * 4, **big),
        world, keep_engines=True, graph=graph)
    rep = []
    for e in engines:
        rep += ssb.records_to_lines(e.drain_committed())
        e.close()
    ok = ok and sorted(rep) == O.repeat_closure(init_lines, new_lines)
    return ok, cfg, dict(init=len(init_lines), new=len(new_lines), repeat=len(rep), rollbacks=sum(x.rollbacks for x in st2))


def one_sc(rng):
    """State-change (SC) what-if queries: a few LPs get new road sets from time 0 on; --diff_repeat on the recorded
    run must give init output minus what the change supersedes plus its new lines == ONE run on the changed network."""
    n_nodes = int(rng.choice([12, 100, 400]))
    deg = int(rng.choice([1, 2, 4]))
    n_trips = int(rng.choice([20, 500, 3000]))
    seed = int(rng.integers(0, 1 << 30))
    bucket = int(rng.choice([5, 11, 64]))
    rbucket = int(rng.choice([5, 11, 64, 400]))
    rwindow = int(rng.choice([1, 2, 8]))
    n_sc = int(rng.choice([1, 2, 5]))
    cfg = dict(kind="sc", n_nodes=n_nodes, deg=deg, n_trips=n_trips, seed=seed, bucket=bucket, rbucket=rbucket, rwindow=rwindow, n_sc=n_sc)
    if os.environ.get("FUZZ_VERBOSE"):
        print("sc", cfg, flush=True)
    sc = _random_network(n_nodes, deg, n_trips, seed)
    r2 = np.random.default_rng(seed + 1)
    lps = r2.choice(n_nodes, size=min(n_sc, n_nodes), replace=False)
    changed = traffic.Scenario(sc.n_lp, sc.state_ids, sc.road_start, sc.road_dst, sc.road_speed.copy(), sc.road_len.copy(),
                               sc.trip_id, sc.trip_dep, sc.route_start, sc.route_nodes)
    queries = []
    for lp in lps:
        a, b = int(sc.road_start[lp]), int(sc.road_start[lp + 1])
        changed.road_speed[a:b] = r2.integers(0, 61, b - a)
        changed.road_len[a:b] = r2.integers(1, 1200, b - a)
        queries.append((int(lp), 0, [int(x) for x in sc.road_dst[a:b]], [int(x) for x in changed.road_speed[a:b]],
                        [int(x) for x in changed.road_len[a:b]]))
    big = dict(max_events=(sc.route_nodes.size + 64) * 12 + (1 << 16), max_batch=sc.route_nodes.size * 4 + (1 << 14))
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=bucket, **big)
    init = set(ssb.records_to_lines(ssb.raw_to_records(rec)))
    o = O.Sim.traffic_scenario(changed)
    o.run_closure(2)
    full = set(O.records_to_lines(o.committed()))
    z = np.zeros(0, dtype=np.int64)
    wi = traffic.WhatIf(z, z, np.zeros(1, dtype=np.int64), z, np.zeros((0, 3), dtype=np.int64), queries)
    eng = traffic.make_repeat_engine(sc, hs, wi, bucket_ticks=rbucket, window_buckets=rwindow, keep_records=True,
                                     max_committed=(sc.route_nodes.size + 64) * 4, **big)
    st2 = eng.run()
    lines = ssb.records_to_lines(eng.drain_committed())
    eng.close()
    new = set(lines) - init
    ok = ((init & full) | new) == full and set(lines) <= full and len(set(lines)) == len(lines)
    if not ok and st2.duplicates > 0:
        # a vehicle that passes the changed junction several times (cycles in the random network): an event of its
        # superseded trajectory and one of the new trajectory can carry the same key (time, id) with different route
        # positions — the key collision of R1 (see one_phold): the reference's merge cannot tell them apart either
        return None, cfg, dict(collision=True)
    return ok, cfg, dict(init=len(init), full=len(full), new=len(new), repeat=len(lines), rollbacks=st2.rollbacks)


class RefHang(Exception):
    pass


def _run_lines(cmd, env=None, tries=4):
    """stdout "RE," lines of a binary, sorted; the reference has an intermittent start-up crash: retry"""
    import subprocess
    for _ in range(tries):
        try:
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=float(os.environ.get("FUZZ_REF_TIMEOUT", "60")),
                               env=dict(os.environ, **(env or {})))
        except subprocess.TimeoutExpired:
            raise RefHang(cmd[0])      # the reference itself sometimes never terminates (seen with --diff_init): not a finding here
        if p.returncode == 0:
            return sorted(l for l in p.stdout.splitlines() if l.startswith("RE,"))
    if "oracle/_ref" in cmd[0]:
        raise RefHang(cmd[0])      # ... or dies on every try (seen with --diff_init on some networks)
    raise RuntimeError(f"{cmd[0]} failed: {p.stderr[-300:]}")


def one_ref(rng):
    """Three-way check on a random input: the UNMODIFIED reference binary (oracle/_ref, built from /root/reference where
    that tree exists) == the oracle's restatement == the drop-in binary on the emulated kernels (tools/emu/apps), line
    for line. Pins the oracle on the reference beyond the committed goldens."""
    import tempfile
    ref_dir, emu_dir = os.path.join(ROOT, "oracle", "_ref"), os.path.join(HERE, "apps")
    window = str(int(rng.choice([1, 2, 5])))
    if rng.random() < 0.5:
        n_lp = int(rng.choice([1, 8, 50, 200]))
        n_init = int(rng.choice([1, 40, 600, 2000]))
        ratio = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
        lam = float(rng.choice([0.5, 1.0, 3.0]))
        cfg = dict(kind="ref-phold", n_lp=n_lp, n_init=n_init, ratio=ratio, lam=lam, window=window)
        if os.environ.get("FUZZ_VERBOSE"):
            print("ref", cfg, flush=True)
        args = [str(n_lp), str(n_init), str(ratio), str(lam), "0", "50"]
        ref = _run_lines([os.path.join(ref_dir, "PHOLD")] + args, {"SCALESIM_NUM_THR": str(int(rng.choice([1, 2, 4])))})
        emu = _run_lines([os.path.join(emu_dir, "PHOLD")] + args, {"SCALESIM_B200_WINDOW": window})
        s = O.Sim.phold(n_lp, n_init, ratio, lam)
        s.run_closure(2)
    else:
        n_nodes = int(rng.choice([4, 30, 150]))
        deg = int(rng.choice([1, 2, 5]))
        n_trips = int(rng.choice([1, 30, 400]))
        seed = int(rng.integers(0, 1 << 30))
        cfg = dict(kind="ref-traffic", n_nodes=n_nodes, deg=deg, n_trips=n_trips, seed=seed, window=window)
        if os.environ.get("FUZZ_VERBOSE"):
            print("ref", cfg, flush=True)
        sc = _random_network(n_nodes, deg, n_trips, seed)
        sc.partition = np.zeros(n_nodes, dtype=np.int64)
        from scalesim_b200 import scenario
        with tempfile.TemporaryDirectory() as td:
            rd, trip, part = (os.path.join(td, n) for n in ("rd.csv", "trip.csv", "part"))
            scenario.write_traffic_csv(sc, rd, trip, part)
            ref = _run_lines([os.path.join(ref_dir, "TrafficSim"), part, rd, trip], {"SCALESIM_NUM_THR": str(int(rng.choice([2, 4])))})
            emu = _run_lines([os.path.join(emu_dir, "TrafficSim"), part, rd, trip], {"SCALESIM_B200_WINDOW": window})
        s = O.Sim.traffic_scenario(sc)
        s.run_closure(2)
    orc = sorted(O.records_to_lines(s.committed()))
    return ref == orc == emu, cfg, dict(ref=len(ref), oracle=len(orc), emu=len(emu))


def _ref_store_keys(path):
    """keys of one stand-in LevelDB file the reference wrote (oracle/shim/leveldb/db.h: [u64 klen][key][u64 vlen][value]...;
    key = 60 bytes lp | time | id, leveldb_store.hpp:336-368)"""
    import struct
    b = open(os.path.join(path, "data.bin"), "rb").read()
    o, out = 0, []
    while o < len(b):
        (a,) = struct.unpack_from("<Q", b, o); o += 8
        k = b[o:o + a]; o += a
        (v,) = struct.unpack_from("<Q", b, o); o += 8 + v
        f = k.rstrip(b"\0").split(b"\0")
        out.append([-int(x.split(b"-")[1]) if b"-" in x else int(x) for x in f])
    return sorted(out)


def one_refstore(rng):
    """--diff_init on a random road network: the three stores the UNMODIFIED reference writes (event / cancel / state
    key sets, for keys below finish_time) == the key sets derived from the history store the drop-in binary writes on
    the emulated kernels; and both print the same lines."""
    import subprocess
    import tempfile
    from scalesim_b200 import scenario, store
    n_nodes = int(rng.choice([4, 30, 150]))
    deg = int(rng.choice([1, 2, 5]))
    n_trips = int(rng.choice([1, 30, 300]))
    seed = int(rng.integers(0, 1 << 30))
    window = str(int(rng.choice([1, 2, 5])))
    cfg = dict(kind="refstore", n_nodes=n_nodes, deg=deg, n_trips=n_trips, seed=seed, window=window)
    if os.environ.get("FUZZ_VERBOSE"):
        print("refstore", cfg, flush=True)
    sc = _random_network(n_nodes, deg, n_trips, seed)
    sc.partition = np.zeros(n_nodes, dtype=np.int64)
    with tempfile.TemporaryDirectory() as td:
        rd, trip, part = (os.path.join(td, n) for n in ("rd.csv", "trip.csv", "part"))
        scenario.write_traffic_csv(sc, rd, trip, part)
        st_ref, st_emu = os.path.join(td, "ref"), os.path.join(td, "emu")
        os.makedirs(st_ref)
        ref = _run_lines([os.path.join(ROOT, "oracle", "_ref", "TrafficSim"), "--sim_id=7", f"--store_dir={st_ref}", "--diff_init", part, rd, trip],
                         {"SCALESIM_NUM_THR": "2"})
        emu = _run_lines([os.path.join(HERE, "apps", "TrafficSim"), "--sim_id=7", f"--store_dir={st_emu}", "--diff_init", part, rd, trip],
                         {"SCALESIM_B200_WINDOW": window})
        fin = traffic.FINISH_TIME
        keys = {db: [k for k in _ref_store_keys(os.path.join(st_ref, "7", db + "0")) if k[1] < fin] for db in ("event", "cancel", "state")}
        hs = store.HistoryStore.load(st_emu, 7, 0)
        # (lp, -1, -1) records of LPs reached only by events at or beyond finish_time depend on the reference's thread
        # timing (dequeued before the loop ends or not): compared on the LPs that executed a committed event
        live = set(int(x) for x in hs.active_lps())
        ref_state = [k for k in keys["state"] if k[1] >= 0 or k[0] in live]
        ok = ref == emu and hs.event_keys().tolist() == keys["event"] and hs.cancel_keys().tolist() == keys["cancel"] \
            and hs.state_keys().tolist() == ref_state
    return ok, cfg, dict(lines=len(ref), events=len(keys["event"]), cancels=len(keys["cancel"]), states=len(keys["state"]))


def one_refrepeat(rng):
    """--diff_init then --diff_repeat with added trips (AE lines) through the UNMODIFIED reference and through the
    drop-in binary on the emulated kernels. The reference's repeat output depends on its thread timing (it can lose
    re-executed history lines, never gain any — tests/golden/make_golden_diff.py): every reference run must be contained
    in the drop-in's output, which must be the closure; the new lines (what the added trips produce) must be equal."""
    import tempfile
    from scalesim_b200 import scenario
    n_nodes = int(rng.choice([6, 30, 120]))
    deg = int(rng.choice([1, 2, 4]))
    n_trips = int(rng.choice([5, 60, 300]))
    n_add = int(rng.choice([1, 3, 10]))
    seed = int(rng.integers(0, 1 << 30))
    window = str(int(rng.choice([1, 2, 5])))
    cfg = dict(kind="refrepeat", n_nodes=n_nodes, deg=deg, n_trips=n_trips, n_add=n_add, seed=seed, window=window)
    if os.environ.get("FUZZ_VERBOSE"):
        print("refrepeat", cfg, flush=True)
    both = _random_network(n_nodes, deg, n_trips + n_add, seed)
    cut = int(both.route_start[n_trips])
    sc = traffic.Scenario(both.n_lp, both.state_ids, both.road_start, both.road_dst, both.road_speed, both.road_len,
                          both.trip_id[:n_trips], both.trip_dep[:n_trips], both.route_start[:n_trips + 1], both.route_nodes[:cut],
                          np.zeros(n_nodes, dtype=np.int64))
    with tempfile.TemporaryDirectory() as td:
        rd, trip, part, ae = (os.path.join(td, n) for n in ("rd.csv", "trip.csv", "part", "ae.csv"))
        scenario.write_traffic_csv(sc, rd, trip, part)
        with open(ae, "w") as f:
            for t in range(n_trips, n_trips + n_add):
                route = ",".join(str(x) for x in both.route_nodes[both.route_start[t]:both.route_start[t + 1]])
                f.write(f"AE,{100000 + t},0,{both.trip_dep[t]},{route}\n")
        st_ref, st_emu = os.path.join(td, "ref"), os.path.join(td, "emu")
        os.makedirs(st_ref)
        flags = lambda st: ["--sim_id=7", f"--store_dir={st}"]
        ref_bin, emu_bin = os.path.join(ROOT, "oracle", "_ref", "TrafficSim"), os.path.join(HERE, "apps", "TrafficSim")
        ref_init = _run_lines([ref_bin] + flags(st_ref) + ["--diff_init", part, rd, trip], {"SCALESIM_NUM_THR": "2"})
        emu_init = _run_lines([emu_bin] + flags(st_emu) + ["--diff_init", part, rd, trip], {"SCALESIM_B200_WINDOW": window})
        refs = [_run_lines([ref_bin] + flags(st_ref) + ["--diff_repeat", part, rd, ae], {"SCALESIM_NUM_THR": str(t)}) for t in (2, 4)]
        emu = _run_lines([emu_bin] + flags(st_emu) + ["--diff_repeat", part, rd, ae], {"SCALESIM_B200_WINDOW": window})
    init = set(emu_init)
    new = sorted(l for l in emu if l not in init)
    ok = ref_init == emu_init and emu == O.repeat_closure(emu_init, new)
    visited = set(int(l.split(",")[4]) for l in emu_init)
    if not set(int(x) for x in both.route_nodes[cut:]) <= visited:
        # An added trip passes a junction that executed nothing in the recorded run: the reference has no state for it
        # in its store (it stores only LPs its scheduler dequeued) and its repeat run then uses another junction's
        # record or none — wrong travel times or missing trajectories. The drop-in's answer is still checked against
        # the closure; the reference's is not comparable.
        return ("skip" if ok else False), cfg, dict(reference_not_comparable="added trip passes a junction without recorded state")
    for r in refs:
        ok = ok and set(r) <= set(emu) and sorted(l for l in r if l not in init) == new
    return ok, cfg, dict(init=len(emu_init), new=len(new), repeat=len(emu), ref_sizes=[len(r) for r in refs])


def one_refphold(rng):
    """PHOLD --diff_init then --diff_repeat with NUM_WHAT_IF added events (phold::init_what_if: id NUM_INIT_MSG + i at LP i,
    time 1; phold.hpp:232-246) through the unmodified reference and the drop-in binary: equal init lines and store key
    sets; every reference repeat run contained in the drop-in's output (= the closure), equal new lines."""
    import tempfile
    from scalesim_b200 import store
    n_lp = int(rng.choice([4, 20, 100]))
    n_init = int(rng.choice([n_lp, 4 * n_lp, 16 * n_lp]))
    ratio = float(rng.choice([0.0, 0.1, 0.5]))
    lam = float(rng.choice([0.5, 1.0, 3.0]))
    n_wi = int(rng.choice([1, 3]))
    window = str(int(rng.choice([1, 2, 5])))
    cfg = dict(kind="refphold", n_lp=n_lp, n_init=n_init, ratio=ratio, lam=lam, n_wi=n_wi, window=window)
    if os.environ.get("FUZZ_VERBOSE"):
        print("refphold", cfg, flush=True)
    args = [str(n_lp), str(n_init), str(ratio), str(lam), str(n_wi), "50"]
    with tempfile.TemporaryDirectory() as td:
        st_ref, st_emu = os.path.join(td, "ref"), os.path.join(td, "emu")
        os.makedirs(st_ref)
        flags = lambda st: ["--sim_id=3", f"--store_dir={st}"]
        ref_bin, emu_bin = os.path.join(ROOT, "oracle", "_ref", "PHOLD"), os.path.join(HERE, "apps", "PHOLD")
        ref_init = _run_lines([ref_bin] + flags(st_ref) + ["--diff_init"] + args, {"SCALESIM_NUM_THR": "2"})
        emu_init = _run_lines([emu_bin] + flags(st_emu) + ["--diff_init"] + args, {"SCALESIM_B200_WINDOW": window})
        fin = phold.FINISH_TIME
        keys = {db: [k for k in _ref_store_keys(os.path.join(st_ref, "3", db + "0")) if k[1] < fin] for db in ("event", "cancel", "state")}
        hs = store.HistoryStore.load(st_emu, 3, 0)
        live = set(int(x) for x in hs.active_lps())
        ok = ref_init == emu_init and hs.event_keys().tolist() == keys["event"] and hs.cancel_keys().tolist() == keys["cancel"] \
            and hs.state_keys().tolist() == [k for k in keys["state"] if k[1] >= 0 or k[0] in live]
        refs = [_run_lines([ref_bin] + flags(st_ref) + ["--diff_repeat"] + args, {"SCALESIM_NUM_THR": str(t)}) for t in (2, 4)]
        emu = _run_lines([emu_bin] + flags(st_emu) + ["--diff_repeat"] + args, {"SCALESIM_B200_WINDOW": window})
    init = set(emu_init)
    new = sorted(l for l in emu if l not in init)
    ok = ok and emu == O.repeat_closure(emu_init, new)
    if not set(range(n_wi)) <= live:
        return ("skip" if ok else False), cfg, dict(reference_not_comparable="a what-if event reaches an LP without recorded state")
    for r in refs:
        ok = ok and set(r) <= set(emu) and sorted(l for l in r if l not in init) == new
    return ok, cfg, dict(init=len(emu_init), new=len(new), repeat=len(emu), ref_sizes=[len(r) for r in refs])


def one_pholddiff(rng):
    """PHOLD --diff_init / --diff_repeat through the C ABI: the history goes back as processed events
    (ssb_load_history), random what-if events are injected (new ids, random LPs and times), and the repeat output must
    be the closure of those events over the recorded output; init + new == one full run with the extra events."""
    n_lp = int(rng.choice([1, 8, 100, 800]))
    n_init = int(rng.choice([20, 400, 4000]))
    ratio = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
    lam = float(rng.choice([0.5, 1.0, 4.0]))
    n_wi = int(rng.choice([1, 3, 20]))
    window, rwindow = int(rng.choice([1, 2, 8])), int(rng.choice([1, 2, 8]))
    cfg = dict(kind="pholddiff", n_lp=n_lp, n_init=n_init, ratio=ratio, lam=lam, n_wi=n_wi, window=window, rwindow=rwindow)
    if os.environ.get("FUZZ_VERBOSE"):
        print("pholddiff", cfg, flush=True)
    tb = tables(ratio, lam, 1)
    kw = dict(tables=tb, max_committed=n_init * 12 + 4096, max_events=n_init * 24 + (1 << 16), max_batch=n_init * 4 + (1 << 14))
    eng = phold.make_engine(n_lp, n_init, ratio, lam, flags=ssb.FLAG_HISTORY, window_buckets=window, **kw)
    eng.run()
    rec, sent = eng.drain_history()
    eng.close()
    init_lines = sorted(ssb.records_to_lines(ssb.raw_to_records(rec)))
    ev = np.zeros(n_wi, dtype=ssb.EVENT_DTYPE)
    ev["id"] = n_init + np.arange(n_wi)
    ev["src"] = ev["dst"] = rng.integers(0, n_lp, n_wi)
    ev["recv_time"] = ev["send_time"] = rng.integers(1, phold.FINISH_TIME, n_wi)
    # one full run with the extra events: the definition of what the repeat run has to add
    full = phold.make_engine(n_lp, n_init, ratio, lam, window_buckets=1, keep_records=True, **kw)
    full.load_events(ev)
    full.run()
    full_lines = sorted(ssb.records_to_lines(full.drain_committed()))
    full.close()
    new_lines = sorted(set(full_lines) - set(init_lines))
    ok = set(init_lines) <= set(full_lines)                      # PHOLD chains do not interact
    eng = phold.make_engine(n_lp, n_init, ratio, lam, load=False, window_buckets=rwindow, **kw)
    eng.load_history(rec, sent)
    eng.load_events(ev)
    st = eng.run()
    rep = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    ok = ok and sorted(set(rep) - set(init_lines)) == new_lines and rep == O.repeat_closure(init_lines, new_lines)
    return ok, cfg, dict(init=len(init_lines), new=len(new_lines), repeat=len(rep), rollbacks=st.rollbacks)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0, n, bad, rb, cap, col, hang, skip = time.time(), 0, 0, 0, 0, 0, 0, 0
    while time.time() - t0 < budget:
        r = rng.random()
        kind = one_phold if r < 0.3 else (one_traffic if r < 0.5 else (one_diff if r < 0.7 else (one_multi if r < 0.82 else (one_multidiff if r < 0.89 else (one_sc if r < 0.95 else one_pholddiff)))))
        if os.environ.get("FUZZ_ONLY"):
            kind = globals()["one_" + os.environ["FUZZ_ONLY"]]
        try:
            ok, cfg, info = kind(rng)
        except RefHang:
            hang += 1
            continue
        except ssb.EngineError as e:
            # a capacity overflow (pool / batch / snapshot ring too small for a hostile setting) must be a clean
            # SSB_ERR_CAPACITY — it counts as handled; anything else is a finding
            ok, cfg, info = e.status == 4, {"kind": kind.__name__}, {"exception": repr(e)}
            cap += int(ok)
        except Exception as e:      # noqa: BLE001
            ok, cfg, info = False, {"kind": kind.__name__}, {"exception": repr(e)}
        n += 1
        rb += int(info.get("rollbacks", 0) > 0)
        if ok is None:
            col += 1
            continue
        if ok == "skip":
            skip += 1
            continue
        if not ok:
            bad += 1
            print("MISMATCH", kind.__name__, cfg, info, flush=True)
    print(f"{n} runs ({rb} with rollbacks, {cap} clean capacity errors, {col} key collisions (R1), {hang} runs the reference binary itself did not survive, {skip} where its answer is not comparable), {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
