"""GPU parity tests (through the C ABI): PHOLD on the sm_100a kernels vs the reference's committed output
(tests/golden, produced by the unmodified reference) and vs the CPU oracle on the same seeded inputs.
Bit-exact: the sorted multiset of "RE,..." lines, or its digest, must be identical."""
import numpy as np
import pytest

import oracle_lib as O
import scalesim_b200 as ssb
from scalesim_b200 import phold
from fixtures import golden, golden_lines

pytestmark = pytest.mark.gpu


def run_phold(tables, n_lp, n_init, ratio, lam, **kw):
    eng = phold.make_engine(n_lp, n_init, ratio, lam, tables=tables, **kw)
    st = eng.run()
    return eng, st


@pytest.mark.parametrize("window", [1, 2, 5, 50])
def test_phold_100_lines_equal_reference(phold_tables_01, window):
    g = golden("phold_100_1600")
    eng, st = run_phold(phold_tables_01, 100, 1600, 0.1, 1.0, window_buckets=window, keep_records=True)
    rec = eng.drain_committed()
    assert st.committed == g["count"] == rec.size
    assert sorted(ssb.records_to_lines(rec)) == golden_lines("phold_100_1600")
    assert eng.digest() == g["digest"]
    if window == 1:
        assert st.rollbacks == 0 and st.antis_sent == 0      # window <= lookahead: provably no straggler
        assert st.processed == st.committed
    if window == 50:
        assert st.rollbacks > 0 and st.antis_sent > 0        # the whole run in one window: heavy optimism
    eng.close()


@pytest.mark.parametrize("name", ["phold_64_1024", "phold_1000_4000"])
@pytest.mark.parametrize("window", [1, 3])
def test_phold_shapes_equal_reference(name, window):
    g = golden(name)
    n_lp, n_init, ratio, lam = g["args"]
    tables = phold.build_tables(ratio, lam)
    eng, st = run_phold(tables, int(n_lp), int(n_init), ratio, lam, window_buckets=window)
    assert eng.digest() == g["digest"]
    eng.close()


@pytest.mark.parametrize("bucket,window", [(10000, 1), (10000, 4), (2500, 2), (25000, 1), (50, 400)])
def test_phold_10000_digest_equal_reference(phold_tables_01, bucket, window):
    """Optimism invariance: the committed result does not depend on the bucket width or the window."""
    g = golden("phold_10000_160000")
    eng, st = run_phold(phold_tables_01, 10000, 160000, 0.1, 1.0, bucket_ticks=bucket, window_buckets=window)
    assert st.committed == 728183
    assert eng.digest() == g["digest"]
    eng.close()


def test_phold_step_by_step_matches_run(phold_tables_01):
    eng = phold.make_engine(100, 1600, 0.1, 1.0, tables=phold_tables_01)
    rounds = 0
    while not eng.step():
        rounds += 1
        assert rounds < 1000
    assert eng.digest() == golden("phold_100_1600")["digest"]
    st = eng.stats()
    assert st.gvt_time >= phold.FINISH_TIME
    eng.close()


@pytest.mark.parametrize("window", [1, 2, 8])
def test_stateful_model_equals_sequential_oracle(phold_tables_01, window):
    """Mutable LP state that feeds back into events: any ordering, rollback or state-restore mistake changes the
    committed records and the final states. Oracle: plain sequential DES."""
    n_lp, n_init = 500, 4000
    s = O.Sim.phold(n_lp, n_init, 0.1, 1.0, stateful=True)
    s.run_sequential()
    want_ids, want_words = s.final_states(n_lp)
    eng, st = run_phold(phold_tables_01, n_lp, n_init, 0.1, 1.0, window_buckets=window, stateful=True,
                        keep_records=True)
    assert eng.digest() == s.digest()
    rec = eng.drain_committed()
    assert sorted(ssb.records_to_lines(rec)) == sorted(O.records_to_lines(s.committed()))
    got = eng.read_states(want_ids)
    assert np.array_equal(got, want_words)
    if window > 1:
        assert st.rollbacks > 0 and st.undone > 0
    eng.close()


def _crafted_engine(window=1):
    """PHOLD kernels with hand-made tables (16 entries, LOOK_AHEAD 0) so that every hop is chosen by the test.
    idx = (id + hops') % 16;  remote entry: dst = latency % NUM_LP."""
    lat = np.full(16, 5000, dtype=np.int64)       # default: local, lands beyond finish_time (dropped, never executed)
    rem = np.zeros(16, dtype=np.int32)
    lat[3], rem[3] = 24, 1                        # id 2, first hop: LP1@5 -> LP0@29
    lat[5], rem[5] = 1, 1                         # id 4, first hop: LP2@1 -> LP1@2
    cfg = ssb.Config(model=ssb.MODEL_PHOLD, n_lp=4, finish_time=1000, bucket_ticks=100, window_buckets=window,
                     max_events=1 << 12, max_batch=1 << 10, max_committed=1 << 12)
    eng = ssb.Engine(cfg)
    eng.set_partition(None)
    eng.set_model_data(0, lat)
    eng.set_model_data(1, rem)
    eng.set_model_data(2, np.array([0, 4], dtype=np.int64))
    ids, words = phold.init_states(4)
    eng.load_states(ids, words)
    return eng


def _ev(i, t, dst, cancel=False):
    e = np.zeros(1, dtype=ssb.EVENT_DTYPE)
    e[0] = (i, dst, dst, t, t, 0, 0, 1 if cancel else 0)
    return e


def test_crafted_straggler_and_antimessage_cascade():
    """Rules R4-R6 on the device, step by step (the device counterpart of logical_process_test.cc:472-784).
    round 1: LP0 runs 10,20,40; LP1 runs id2@5 (-> LP0@29); LP2 runs id4@1 (-> LP1@2)
    round 2: LP0 gets straggler 29 < 40: undo 40 (+ its anti), run 29, 40 again.
             LP1 gets straggler 2 < 5 : undo 5, anti-message for (29,id2) to LP0, run 2, run 5 again (re-sends 29)
    round 3: LP0 gets anti(29,id2) then the re-sent (29,id2): undo 29 and 40, annihilate the old 29, run the new 29, 40."""
    eng = _crafted_engine()
    eng.load_events(np.concatenate([_ev(0, 10, 0), _ev(1, 20, 0), _ev(3, 40, 0), _ev(2, 5, 1), _ev(4, 1, 2)]))
    assert not eng.step()
    st = eng.stats()
    assert (st.processed, st.rollbacks, st.undone, st.antis_sent) == (5, 0, 0, 0)
    assert not eng.step()
    st = eng.stats()
    assert (st.processed, st.rollbacks, st.undone, st.antis_sent) == (5 + 4, 2, 2, 2)
    assert st.annihilated == 0
    assert not eng.step()
    st = eng.stats()
    assert (st.processed, st.rollbacks, st.undone) == (9 + 2, 3, 4)
    assert st.annihilated == 1 and st.duplicates == 0
    assert st.antis_sent == 2 + 2
    done = False
    for _ in range(4):
        done = eng.step()
        if done:
            break
    assert done
    rec = eng.drain_committed()
    got = sorted((int(r["dst"]), int(r["recv_time"]), int(r["id"]), int(r["src"]), int(r["send_time"])) for r in rec)
    assert got == [(0, 10, 0, 0, 10), (0, 20, 1, 0, 20), (0, 29, 2, 1, 5), (0, 40, 3, 0, 40),
                   (1, 2, 4, 2, 1), (1, 5, 2, 1, 5), (2, 1, 4, 2, 1)]
    assert eng.stats().committed == 7
    eng.close()


def test_crafted_duplicate_and_unmatched_anti():
    """queue.hpp:92 (second insert of a key ignored), queue.hpp:87-90 (anti without a match is dropped; anti in the
    same batch erases the positive: logical_process_test.cc:142-198)."""
    eng = _crafted_engine()
    eng.load_events(np.concatenate([
        _ev(0, 10, 0), _ev(0, 10, 0),                         # duplicate key: one survives
        _ev(7, 15, 0),
        _ev(9, 50, 3, cancel=True),                           # unmatched anti-message: dropped
    ]))
    # ev, cancel, ev -> exactly one survives. Arrival order matters here, and it is only defined across calls
    # (inside one ssb_load_events call equal keys may be appended in any order).
    eng.load_events(_ev(7, 15, 0, cancel=True))          # annihilated inside the batch
    eng.load_events(_ev(8, 16, 0))
    eng.load_events(_ev(8, 16, 0, cancel=True))
    eng.load_events(_ev(8, 16, 0))
    eng.run()
    st = eng.stats()
    assert st.duplicates == 1 and st.annihilated == 2
    rec = eng.drain_committed()
    assert sorted((int(r["dst"]), int(r["recv_time"]), int(r["id"])) for r in rec) == [(0, 10, 0), (0, 16, 8)]
    eng.close()


def test_duplicates_in_a_crowded_hash_set():
    """R1 (queue.hpp:92) when the round's fingerprint set is as full as it gets (max_batch = the number of arrivals,
    so the set has 2 slots per arrival): most probe sequences collide, their tails are walked at the end of k_exec
    (parked in shared memory), and every duplicate must still be found there. Every event lands beyond finish_time
    after one hop, so the committed set is exactly the set of distinct (LP, time, id) triples that were loaded."""
    n_lp = 64
    lat = np.full(16, 5000, dtype=np.int64)
    rem = np.zeros(16, dtype=np.int32)
    cfg = ssb.Config(model=ssb.MODEL_PHOLD, n_lp=n_lp, finish_time=1000, bucket_ticks=100, window_buckets=1,
                     max_events=1 << 13, max_batch=1024, max_committed=1 << 12)
    eng = ssb.Engine(cfg)
    eng.set_partition(None)
    eng.set_model_data(0, lat)
    eng.set_model_data(1, rem)
    eng.set_model_data(2, np.array([0, n_lp], dtype=np.int64))
    ids, words = phold.init_states(n_lp)
    eng.load_states(ids, words)
    rng = np.random.default_rng(11)
    uniq = set()
    while len(uniq) < 900:
        uniq.add((int(rng.integers(0, n_lp)), int(rng.integers(0, 100)), int(rng.integers(0, 50))))
    uniq = sorted(uniq)
    dup = [uniq[i] for i in rng.choice(len(uniq), 124, replace=False)]
    allev = uniq + dup
    ev = np.zeros(len(allev), dtype=ssb.EVENT_DTYPE)
    for k, (dst, t, i) in enumerate(allev):
        ev[k] = (i, dst, dst, t, t, 0, 0, 0)
    eng.load_events(ev[rng.permutation(len(allev))])
    eng.run()
    st = eng.stats()
    assert st.duplicates == 124 and st.committed == 900
    rec = eng.drain_committed()
    assert sorted((int(r["dst"]), int(r["recv_time"]), int(r["id"])) for r in rec) == uniq
    eng.close()


@pytest.mark.parametrize("window", [1, 10])
def test_crafted_scenario_result_is_window_invariant(window):
    eng = _crafted_engine(window)
    eng.load_events(np.concatenate([_ev(0, 10, 0), _ev(1, 20, 0), _ev(3, 40, 0), _ev(2, 5, 1), _ev(4, 1, 2)]))
    eng.run()
    assert eng.stats().committed == 7
    eng.close()


def test_range_and_capacity_errors(phold_tables_01):
    lat, rem = phold_tables_01
    cfg = ssb.Config(model=ssb.MODEL_PHOLD, n_lp=4, finish_time=1000, bucket_ticks=100, max_events=1 << 10,
                     max_batch=1 << 10)
    eng = ssb.Engine(cfg)
    bad = np.zeros(1, dtype=ssb.EVENT_DTYPE)
    bad[0] = (1, 0, 7, 10, 10, 0, 0, 0)      # destination >= n_lp
    with pytest.raises(ssb.EngineError) as ei:
        eng.load_events(bad)
    assert ei.value.status == 3              # SSB_ERR_RANGE
    eng.close()
    eng = ssb.Engine(cfg)
    with pytest.raises(ssb.EngineError) as ei:
        eng.step()                           # model data not set
    assert ei.value.status == 5              # SSB_ERR_MODEL
    eng.close()


def test_phold_1m_full_size_count_and_digest(phold_tables_01):
    """BASELINE.json config 2 at full size: 72 823 220 committed and the digest of the oracle's
    closure. Also the size-independent property: count == processed when the window is within the lookahead."""
    n_lp, n_init = 1_000_000, 16_000_000
    eng, st = run_phold(phold_tables_01, n_lp, n_init, 0.1, 1.0)
    assert st.committed == 72_823_220
    assert st.processed == st.committed and st.rollbacks == 0
    s = O.Sim.phold(n_lp, n_init, 0.1, 1.0)
    s.run_closure(8, keep_records=False)
    assert eng.digest() == s.digest()
    eng.close()


@pytest.mark.parametrize("window", [1, 3])
def test_graph_loop_equals_stepwise_loop(phold_tables_01, window):
    """ssb_run (the whole window loop as one CUDA graph with WHILE / IF nodes) and ssb_step (the same kernels launched
    round by round on the stream) must commit the same records with the same statistics."""
    from scalesim_b200 import phold
    eng = phold.make_engine(10000, 160000, 0.1, 1.0, tables=phold_tables_01, window_buckets=window)
    st_graph = eng.run()
    d_graph = eng.digest()
    eng.close()
    eng = phold.make_engine(10000, 160000, 0.1, 1.0, tables=phold_tables_01, window_buckets=window)
    while not eng.step():
        pass
    st_step = eng.stats()
    assert eng.digest() == d_graph == golden("phold_10000_160000")["digest"]
    assert (st_step.committed, st_step.processed, st_step.rollbacks, st_step.antis_sent) == \
           (st_graph.committed, st_graph.processed, st_graph.rollbacks, st_graph.antis_sent)
    eng.close()


def test_compact_load_and_packed_drain_equal_the_wide_forms(phold_tables_01):
    """ssb_load_events_raw (32 B per event) and ssb_drain_committed_packed (24 B per record) carry the same information
    as ssb_load_events / ssb_drain_committed."""
    from scalesim_b200 import phold
    lat = phold_tables_01[0]
    ev = phold.init_events(1000, 16000, lat)
    eng = phold.make_engine(1000, 16000, 0.1, 1.0, tables=phold_tables_01, keep_records=True, load=False)
    eng.load_events(ev)
    eng.run()
    want_digest = eng.digest()
    want = ssb.records_to_lines(eng.drain_committed())
    eng.reset()
    eng.load_events_raw(ssb.events_to_raw(ev))
    eng.run()
    assert eng.digest() == want_digest
    got = ssb.records_to_lines(ssb.packed_to_records(eng.drain_committed_packed()))
    assert sorted(got) == sorted(want)
    # ... and streamed out while the loop runs
    eng.reset()
    eng.load_events_raw(ssb.events_to_raw(ev))
    st, rec = eng.run_drain_packed(100000)
    assert st.committed == rec.size == len(want) and eng.digest() == want_digest
    assert sorted(ssb.records_to_lines(ssb.packed_to_records(rec))) == sorted(want)
    with pytest.raises(ssb.EngineError):
        eng.reset()
        eng.load_events_raw(ssb.events_to_raw(ev))
        eng.run_drain_packed(10)          # output buffer too small: fails loudly
    # ... and through the bounded staging buffers of the streaming form
    eng.reset()
    eng.load_events_raw(ssb.events_to_raw(ev))
    chunks = []
    st = eng.run_stream_packed(lambda rec: chunks.append(rec.copy()))
    rec = np.concatenate(chunks)
    assert st.committed == rec.size == len(want) and eng.digest() == want_digest
    assert sorted(ssb.records_to_lines(ssb.packed_to_records(rec))) == sorted(want)
    eng.close()


def test_committed_log_overflow_fails_loudly(phold_tables_01):
    """A committed log that is too small must not yield a silently shorter output: every drain call fails."""
    from scalesim_b200 import phold
    eng = phold.make_engine(1000, 16000, 0.1, 1.0, tables=phold_tables_01, max_committed=1000)
    st = eng.run()
    assert st.log_dropped == st.committed - 1000 > 0
    for call in (eng.drain_committed, eng.drain_committed_raw, eng.drain_committed_packed):
        with pytest.raises(ssb.EngineError) as ei:
            call()
        assert ei.value.status == 4      # SSB_ERR_CAPACITY
    eng.close()
