"""Known-answer tests of the reference's own LP test-suite (test/medium/logical_process_test.cc), replayed on the
oracle's restatement of lp<App> / eventq / stateq. Line numbers refer to that file."""
import oracle_lib as O

MAX = (2**63 - 1, 2**63 - 1)


def test_insert_event():                      # :49-67
    lp = O.LP()
    lp.buffer(0, 10)
    lp.flush()
    assert lp.dequeue()["recv_time"] == 10
    assert lp.dequeue() is None


def test_direct_insert():                     # :69-99
    lp = O.LP(99)
    lp.buffer(8888, 2, send=1, dst=99)
    lp.buffer(7777, 4, send=3, dst=99)
    lp.flush()
    assert lp.local_time()[0] == 2
    lp.dequeue()
    assert lp.local_time()[0] == 4
    lp.direct_insert(0, 3, send=0, dst=99)
    assert lp.local_time()[0] == 3
    assert lp.dequeue()["id"] == 0
    assert lp.local_time()[0] == 4


def test_dequeue_null_ptr():                  # :101-111
    lp = O.LP()
    assert lp.dequeue() is None
    assert lp.local_time() == MAX


def test_annihilate_inserted_event():         # :113-140
    lp = O.LP()
    lp.buffer(0, 10, send=0)
    lp.buffer(1, 11, send=1)
    lp.flush()
    lp.buffer(0, 10, send=0, cancel=True)
    lp.flush()
    assert lp.dequeue()["id"] == 1


def test_annihilate_buffered_event():         # :142-167
    lp = O.LP()
    lp.buffer(0, 10, send=0)
    lp.buffer(1, 11, send=1)
    lp.buffer(0, 10, send=0, cancel=True)
    lp.flush()
    assert lp.dequeue()["id"] == 1


def test_buffer_double_events_single_cancel():   # :169-198
    lp = O.LP()
    lp.buffer(0, 10)
    lp.buffer(0, 10, cancel=True)
    lp.buffer(0, 10)
    lp.flush()
    r0 = lp.dequeue()
    assert r0 is not None and r0["id"] == 0
    assert lp.dequeue() is None


def test_zero_lookahead_order():              # :200-312: equal receive times are ordered by id
    lp = O.LP()
    for i in (7, 3, 9, 0, 5, 1, 8, 2, 6, 4):
        lp.buffer(i, 10, send=10)
    assert len(lp.flush()) == 0               # :263-284 merging non-stragglers produces no anti-messages
    assert [lp.dequeue()["id"] for _ in range(10)] == list(range(10))
    assert lp.dequeue() is None


def test_set_zero_lookahead_cancel_event():   # :319-347
    lp = O.LP()
    lp.set_cancel(0, 10, 10)
    lp.set_cancel(1, 10, 10)
    lp.buffer(0, 5, send=0)
    c = lp.flush()
    assert len(c) == 2
    assert [int(x["id"]) for x in c] == [0, 1]
    assert [int(x["send_time"]) for x in c] == [10, 10]


def test_cancel_one_of_zero_lookahead_events():   # :349-470: annihilation matches (receive_time, id) exactly
    for order in ("buffered", "inserted"):
        lp = O.LP()
        lp.buffer(0, 10, send=10)
        lp.buffer(1, 10, send=10)
        lp.buffer(2, 10, send=10)
        if order == "inserted":
            lp.flush()
        lp.buffer(1, 10, send=10, cancel=True)
        lp.flush()
        assert [lp.dequeue()["id"] for _ in range(2)] == [0, 2]
        assert lp.dequeue() is None


def _process(lp, ev_id, t):
    r = lp.dequeue()
    assert r["id"] == ev_id
    lp.set_cancel(ev_id, t, t)
    lp.update_state(1, t, ev_id)


def test_rollback_by_event():                 # :472-551
    lp = O.LP()
    for i in (0, 1, 3):
        lp.buffer(i, i, send=i)
    lp.flush()
    for i in (0, 1, 3):
        _process(lp, i, i)
    lp.buffer(2, 2, send=2)
    c = lp.flush()
    assert lp.dequeue()["id"] == 2
    assert lp.dequeue()["id"] == 3
    assert len(c) == 1 and c[0]["id"] == 3


def test_rollback_by_cancel_event():          # :553-625
    lp = O.LP()
    for i in (0, 1, 3):
        lp.buffer(i, i, send=i)
    lp.flush()
    for i in (0, 1, 3):
        _process(lp, i, i)
    lp.buffer(1, 1, send=1, cancel=True)
    c = lp.flush()
    assert len(c) == 2 and c[0]["id"] == 1
    assert lp.dequeue()["id"] == 3


def test_zero_lookahead_rollback_by_event():  # :627-700: same with equal times, ordered by id
    lp = O.LP()
    for i in (0, 1, 3):
        lp.buffer(i, 10, send=10)
    lp.flush()
    for i in (0, 1, 3):
        _process(lp, i, 10)
    lp.buffer(2, 10, send=10)
    c = lp.flush()
    assert [int(x["id"]) for x in c] == [3]
    assert lp.dequeue()["id"] == 2
    assert lp.dequeue()["id"] == 3


def test_zero_lookahead_rollback_by_cancel():  # :702-784: yields 2 anti-messages (ids 2, 3)
    lp = O.LP()
    for i in (0, 1, 2, 3):
        lp.buffer(i, 10, send=10)
    lp.flush()
    for i in (0, 1, 2, 3):
        _process(lp, i, 10)
    lp.buffer(2, 10, send=10, cancel=True)
    c = lp.flush()
    assert [int(x["id"]) for x in c] == [2, 3]
    assert lp.dequeue()["id"] == 3


def test_state_dequeue_update():              # :786-815
    lp = O.LP()
    lp.init_state(100)
    assert lp.get_state() == 100
    lp.update_state(101, 5, 0)
    assert lp.get_state() == 101


def test_state_rollback():                    # :817-980
    for by_cancel in (False, True):
        lp = O.LP()
        lp.init_state(100)
        for i in (0, 1, 3):
            lp.buffer(i, i, send=i)
        lp.flush()
        for i in (0, 1, 3):
            r = lp.dequeue()
            lp.set_cancel(i, i, i)
            lp.update_state(200 + i, i, i)
        assert lp.get_state() == 203
        if by_cancel:
            lp.buffer(1, 1, send=1, cancel=True)
            lp.flush()
            assert lp.get_state() == 200      # versions with key >= (1,1) dropped
        else:
            lp.buffer(2, 2, send=2)
            lp.flush()
            assert lp.get_state() == 201      # versions with key >= (2,2) dropped


def test_commit_and_fossil():                 # queue.hpp:159-177, 203-211
    lp = O.LP()
    for i in range(5):
        lp.buffer(i, 10 * i, send=10 * i)
    lp.flush()
    for i in range(5):
        lp.dequeue()
    out = lp.std_out(25, 0)
    assert [int(x["id"]) for x in out] == [0, 1, 2]
    assert len(lp.std_out(25, 0)) == 0        # exactly once
    lp.clear(25, 0)
    assert lp.size_ev() == 2


def test_timestamp_order():                   # test/small/util_test.cc:15-38: time first, then id
    lp = O.LP()
    lp.buffer(5, 1)
    lp.buffer(1, 2)
    lp.buffer(3, 1)
    lp.flush()
    assert [(int(r["recv_time"]), int(r["id"])) for r in (lp.dequeue(), lp.dequeue(), lp.dequeue())] == \
        [(1, 3), (1, 5), (2, 1)]
