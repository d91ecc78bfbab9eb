"""Device known-answer tests of the Time Warp semantics: the cases of the reference's own LP test-suite
(test/medium/logical_process_test.cc — line numbers refer to that file; the same cases are replayed on the oracle in
tests/test_oracle_kat.py) as ONE-LP scenarios on the kernels, round by round through ssb_step.

The reference cases drive lp<App> directly (buffer / flush_buf / dequeue_event / get_state). The device has no such
per-call surface, so every case is expressed by what those calls make observable:
  * the ORDER in which the LP executes its events: the stateful PHOLD twin folds every executed key into the LP state
    (count' = count + 1, chain' = mix(chain, key)), so the final state equals the fold over the expected dequeue
    sequence only if the order, every rollback and every state restore was right;
  * the statistics of the engine (processed / rollbacks / undone / anti-messages / annihilated / duplicates);
  * the committed records.
Every hop lands beyond finish_time (latency 5000 > 1000), so an event has no descendants; its sent-log entry still
exists and is cancelled by a rollback, exactly like the reference's can_map_ entry."""
import numpy as np
import pytest

import scalesim_b200 as ssb
from scalesim_b200 import phold

pytestmark = pytest.mark.gpu
M64 = (1 << 64) - 1


def _mix64(z):
    z ^= z >> 30; z = (z * 0xbf58476d1ce4e5b9) & M64
    z ^= z >> 27; z = (z * 0x94d049bb133111eb) & M64
    return z ^ (z >> 31)


def fold(keys, lp=0):
    """state {count, chain} of PholdStatefulModel after executing `keys` = [(time, id), ...] in this order"""
    count, chain = 0, lp
    for t, i in keys:
        count += 1
        chain = _mix64(((chain << 32) & M64) ^ ((t << 32) | i)) & 0xFFFFFFFF
    return [count, chain]


def make_lp(window=1, max_batch=1 << 10, n_lp=4):
    lat = np.full(16, 5000, dtype=np.int64)
    rem = np.zeros(16, dtype=np.int32)
    cfg = ssb.Config(model=ssb.MODEL_PHOLD_STATEFUL, n_lp=n_lp, finish_time=1000, bucket_ticks=100, window_buckets=window,
                     max_events=1 << 16, max_batch=max_batch, max_committed=1 << 16)
    eng = ssb.Engine(cfg)
    eng.set_partition(None)
    eng.set_model_data(0, lat)
    eng.set_model_data(1, rem)
    eng.set_model_data(2, np.array([0, n_lp], dtype=np.int64))
    ids, words = phold.init_states(n_lp, stateful=True)
    eng.load_states(ids, words)
    return eng


def ev(i, t, lp=0, cancel=False):
    e = np.zeros(1, dtype=ssb.EVENT_DTYPE)
    e[0] = (i, lp, lp, t, t, 0, 0, 1 if cancel else 0)
    return e


def finish(eng):
    for _ in range(64):
        if eng.step():
            return
    raise AssertionError("simulation did not finish")


def committed_keys(eng):
    rec = eng.drain_committed()
    return sorted((int(r["recv_time"]), int(r["id"])) for r in rec)


def state_of(eng, lp=0):
    return eng.read_states([lp])[0].tolist()


def test_insert_event_and_empty_lp():                     # :49-67, :101-111
    eng = make_lp()
    eng.load_events(ev(0, 10))
    finish(eng)
    assert committed_keys(eng) == [(10, 0)] and state_of(eng) == fold([(10, 0)])
    assert state_of(eng, 1) == [0, 1]                       # an LP without events: nothing dequeued, state untouched
    eng.close()
    eng = make_lp()
    finish(eng)                                             # empty simulation: local time = max, GVT jumps to the end
    assert eng.stats().committed == 0 and eng.stats().processed == 0
    eng.close()


def test_earlier_event_lowers_local_time():               # :69-99 direct_insert: 2, 4 processed... then 3 arrives
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(8888, 2), ev(7777, 4)]))
    assert not eng.step()
    assert eng.stats().processed == 2
    eng.load_events(ev(0, 3))                               # earlier than the processed 4: the LP goes back to it
    assert not eng.step()
    st = eng.stats()
    assert (st.rollbacks, st.undone, st.processed) == (1, 1, 4)
    finish(eng)
    assert state_of(eng) == fold([(2, 8888), (3, 0), (4, 7777)])
    eng.close()


def test_annihilate_inserted_event():                     # :113-140: the positive is already merged (and executed)
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 10), ev(1, 11)]))
    assert not eng.step()
    eng.load_events(ev(0, 10, cancel=True))
    assert not eng.step()
    st = eng.stats()
    assert (st.annihilated, st.rollbacks, st.undone) == (1, 1, 2)
    finish(eng)
    assert committed_keys(eng) == [(11, 1)] and state_of(eng) == fold([(11, 1)])
    eng.close()


def test_annihilate_buffered_event():                     # :142-167: positive and anti-message in the same inbox batch
    eng = make_lp()
    eng.load_events(np.concatenate([ev(0, 10), ev(1, 11)]))
    eng.load_events(ev(0, 10, cancel=True))
    finish(eng)
    st = eng.stats()
    assert (st.annihilated, st.rollbacks, st.processed) == (1, 0, 1)
    assert committed_keys(eng) == [(11, 1)] and state_of(eng) == fold([(11, 1)])
    eng.close()


def test_buffer_double_events_single_cancel():            # :169-198: ev, cancel, ev -> exactly one survives
    eng = make_lp()
    eng.load_events(ev(0, 10))
    eng.load_events(ev(0, 10, cancel=True))
    eng.load_events(ev(0, 10))
    finish(eng)
    st = eng.stats()
    assert (st.annihilated, st.duplicates, st.processed) == (1, 0, 1)
    assert committed_keys(eng) == [(10, 0)] and state_of(eng) == fold([(10, 0)])
    eng.close()


def test_zero_lookahead_events_run_in_id_order():         # :200-312: equal receive times are ordered by id
    eng = make_lp()
    for i in (7, 3, 9, 0, 5, 1, 8, 2, 6, 4):
        eng.load_events(ev(i, 10))
    finish(eng)
    st = eng.stats()
    assert st.antis_sent == 0 and st.rollbacks == 0         # :263-284 merging non-stragglers produces no anti-messages
    assert state_of(eng) == fold([(10, i) for i in range(10)])
    eng.close()


def test_straggler_releases_the_sent_log_from_its_key_on():   # :319-347 set_zero_lookahead_cancel_event
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 10), ev(1, 10)]))    # two processed events at time 10 = two sent-log entries
    assert not eng.step()
    eng.load_events(ev(2, 5))                                   # straggler at 5 releases both, in id order
    assert not eng.step()
    st = eng.stats()
    assert (st.rollbacks, st.undone, st.antis_sent) == (1, 2, 2)
    finish(eng)
    assert state_of(eng) == fold([(5, 2), (10, 0), (10, 1)])
    eng.close()


@pytest.mark.parametrize("anti_first", [False, True])
def test_cancel_one_of_zero_lookahead_events(anti_first):  # :349-470: the match is exact on (receive_time, id)
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 10), ev(1, 10), ev(2, 10)]))
    if not anti_first:
        assert not eng.step()                               # the three are processed before the anti-message arrives
    eng.load_events(ev(1, 10, cancel=True))
    finish(eng)
    st = eng.stats()
    assert st.annihilated == 1
    assert committed_keys(eng) == [(10, 0), (10, 2)] and state_of(eng) == fold([(10, 0), (10, 2)])
    eng.close()


def test_rollback_by_event():                             # :472-551: straggler 2 after 0, 1, 3 -> reprocess 2, 3
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 0), ev(1, 1), ev(3, 3)]))
    assert not eng.step()
    assert eng.stats().processed == 3
    eng.load_events(ev(2, 2))
    assert not eng.step()
    st = eng.stats()
    assert (st.rollbacks, st.undone, st.antis_sent, st.processed) == (1, 1, 1, 5)    # exactly one anti-message (3's output)
    finish(eng)
    assert committed_keys(eng) == [(0, 0), (1, 1), (2, 2), (3, 3)]
    assert state_of(eng) == fold([(0, 0), (1, 1), (2, 2), (3, 3)])
    eng.close()


def test_rollback_by_cancel_event():                      # :553-625: anti-message for the processed event 1
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 0), ev(1, 1), ev(3, 3)]))
    assert not eng.step()
    eng.load_events(ev(1, 1, cancel=True))
    assert not eng.step()
    st = eng.stats()
    assert (st.rollbacks, st.undone, st.antis_sent, st.annihilated) == (1, 2, 2, 1)   # outputs of 1 and 3 are cancelled
    assert st.processed == 3 + 1                                                        # the next dequeue is 3 again
    finish(eng)
    assert committed_keys(eng) == [(0, 0), (3, 3)] and state_of(eng) == fold([(0, 0), (3, 3)])
    eng.close()


def test_zero_lookahead_rollback_by_event():              # :627-700: the same with equal times (order by id)
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 7), ev(1, 7), ev(3, 7)]))
    assert not eng.step()
    eng.load_events(ev(2, 7))
    assert not eng.step()
    st = eng.stats()
    assert (st.rollbacks, st.undone, st.antis_sent) == (1, 1, 1)
    finish(eng)
    assert state_of(eng) == fold([(7, 0), (7, 1), (7, 2), (7, 3)])
    eng.close()


def test_zero_lookahead_rollback_by_cancel():             # :702-784: by cancel: two anti-messages (ids 2 and 3)
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 7), ev(1, 7), ev(2, 7), ev(3, 7)]))
    assert not eng.step()
    eng.load_events(ev(2, 7, cancel=True))
    assert not eng.step()
    st = eng.stats()
    assert (st.rollbacks, st.undone, st.antis_sent, st.annihilated) == (1, 2, 2, 1)
    finish(eng)
    assert committed_keys(eng) == [(7, 0), (7, 1), (7, 3)] and state_of(eng) == fold([(7, 0), (7, 1), (7, 3)])
    eng.close()


def test_state_is_the_most_recently_pushed_one():         # :786-815 state_dequeue_update
    eng = make_lp()
    eng.load_events(np.concatenate([ev(5, 1), ev(6, 2), ev(7, 3)]))
    finish(eng)
    assert state_of(eng) == fold([(1, 5), (2, 6), (3, 7)])
    eng.close()


@pytest.mark.parametrize("by_cancel", [False, True])
def test_state_rollback(by_cancel):                       # :817-980: versions at or above the straggler key are dropped
    eng = make_lp(window=10)
    eng.load_events(np.concatenate([ev(0, 10), ev(1, 20), ev(2, 30), ev(3, 40)]))
    assert not eng.step()
    assert state_of(eng) == fold([(10, 0), (20, 1), (30, 2), (40, 3)])
    eng.load_events(ev(2, 30, cancel=True) if by_cancel else ev(9, 25))
    assert not eng.step()
    want = [(10, 0), (20, 1), (40, 3)] if by_cancel else [(10, 0), (20, 1), (25, 9), (30, 2), (40, 3)]
    assert state_of(eng) == fold(want)                      # restored to the version before the key, then re-executed
    assert eng.stats().undone == 2
    finish(eng)
    assert committed_keys(eng) == want
    eng.close()


def test_many_buffered_events_dequeue_in_global_order():  # :992-1026: 100 threads x 1000 buffers (here: 20 000 events)
    rng = np.random.default_rng(3)
    n = 20000
    times = rng.integers(0, 1000, n)
    order = rng.permutation(n)
    batch = np.zeros(n, dtype=ssb.EVENT_DTYPE)
    batch["id"] = np.arange(n)
    batch["src"] = batch["dst"] = 0
    batch["recv_time"] = batch["send_time"] = times
    eng = make_lp(window=10, max_batch=1 << 15)
    for part in np.array_split(batch[order], 7):
        eng.load_events(part)
    finish(eng)
    keys = sorted(zip(times.tolist(), range(n)))
    assert eng.stats().committed == n
    assert state_of(eng) == fold(keys)
    eng.close()


def test_lp_lookup_by_partition():                        # :1043-1057 lp_mngr::get_lp: id -> LP through the partition
    eng = make_lp(n_lp=7)
    eng.load_events(np.concatenate([ev(0, 10, lp=5), ev(1, 11, lp=2)]))
    finish(eng)
    assert state_of(eng, 5) == fold([(10, 0)], lp=5) and state_of(eng, 2) == fold([(11, 1)], lp=2)
    assert state_of(eng, 0) == [0, 0]
    eng.close()
