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


@pytest.mark.parametrize("bucket,window", [(10000, 1), (10000, 4), (2500, 2), (25000, 1), (7, 3000)])
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


def test_injected_straggler_and_antimessage(phold_tables_01):
    """LP-level rules R4-R6 on the device: a straggler rolls the LP back, an anti-message annihilates a processed
    event, a duplicate key is ignored, an unmatched anti-message is dropped."""
    lat, rem = phold_tables_01
    cfg = ssb.Config(model=ssb.MODEL_PHOLD, n_lp=4, finish_time=1000, bucket_ticks=1000, window_buckets=1,
                     max_events=1 << 12, max_batch=1 << 10, max_committed=1 << 12)
    eng = ssb.Engine(cfg)
    eng.set_partition(None)
    eng.set_model_data(0, lat)
    eng.set_model_data(1, rem)
    eng.set_model_data(2, np.array([10000, 4], dtype=np.int64))
    ids, words = phold.init_states(4)
    eng.load_states(ids, words)

    def ev(i, t, dst=0, cancel=False):
        e = np.zeros(1, dtype=ssb.EVENT_DTYPE)
        e[0] = (i, dst, dst, t, t, 0, 0, 1 if cancel else 0)
        return e

    eng.load_events(np.concatenate([ev(0, 10), ev(1, 20), ev(3, 40)]))
    assert not eng.step()                   # processes 0, 1, 3 (outputs land beyond finish)
    st = eng.stats()
    assert st.processed == 3 and st.rollbacks == 0
    eng.load_events(ev(2, 30))              # straggler
    eng.step()
    st = eng.stats()
    assert st.rollbacks == 1 and st.undone == 1 and st.antis_sent == 1 and st.processed == 5
    eng.load_events(np.concatenate([ev(1, 20, cancel=True), ev(3, 40), ev(9, 50, cancel=True)]))
    eng.step()                              # anti for processed 1 -> rollback to 1; duplicate of 3; unmatched anti
    st = eng.stats()
    assert st.annihilated == 1 and st.duplicates == 1
    assert st.undone == 1 + 3 and st.antis_sent == 1 + 3
    for _ in range(4):
        if eng.step():
            break
    rec = eng.drain_committed()
    assert sorted(int(r["id"]) for r in rec) == [0, 2, 3]
    assert eng.stats().committed == 3
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
