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


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0, n, bad, rb, cap, col = time.time(), 0, 0, 0, 0, 0
    while time.time() - t0 < budget:
        kind = one_phold if rng.random() < 0.6 else one_traffic
        try:
            ok, cfg, info = kind(rng)
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
        if not ok:
            bad += 1
            print("MISMATCH", kind.__name__, cfg, info, flush=True)
    print(f"{n} runs ({rb} with rollbacks, {cap} clean capacity errors, {col} key collisions of the state-mutating model), {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
