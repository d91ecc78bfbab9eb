"""CPU: the exact-differential restatement (oracle_lib.repeat_closure) and the host-side store against the reference's
own --diff_init / --diff_repeat runs (tests/golden/{ring,grid}_diff.json, made by make_golden_diff.py)."""
import numpy as np
import pytest

import oracle_lib as O
from fixtures import golden
from scalesim_b200 import store


@pytest.mark.parametrize("name", ["ring_diff", "grid_diff", "phold_diff"])
def test_repeat_closure_equals_reference_repeat_output(name):
    g = golden(name)
    got = O.repeat_closure(g["init_lines"], g["new_lines"])
    # the reference's undisturbed repeat output; some of its 4-thread runs lose lines (an anti-message overtaking the
    # re-sent positive), never gain any: the union over the recorded runs is the complete answer
    assert got == g["repeat_union"]
    assert set(g["repeat_core"]) <= set(got)
    assert set(g["new_lines"]) <= set(got)


def _deletes(g):
    out = []
    for l in g["de_lines"]:
        f = l.split(",")
        out.append((int(f[1]), int(f[2]), int(f[3])) if f[0] == "DE" else (int(f[4]), int(f[5]), int(f[1])))
    return out


@pytest.mark.parametrize("name", ["ring_diff", "grid_diff"])
def test_delete_closure_equals_reference_repeat_output(name):
    """DE / RE queries: the event is erased, its output is not cancelled (queue.hpp:227-241); on the grid one of the three
    deleted events comes back because another query's cascade re-sends it."""
    g = golden(name)
    got = O.repeat_closure(g["init_lines"], [], _deletes(g))
    assert got == g["delete_union"] == g["delete_core"]


def test_ring_repeat_is_the_19_lines_of_the_survey():
    g = golden("ring_diff")
    assert len(g["repeat_union"]) == 19 and g["repeat_core"] == g["repeat_union"]
    assert "RE,9990,-1,20,0,20,START" in g["repeat_union"] and "RE,0,9,370,0,411,END" in g["repeat_union"]
    assert {k: len(v) for k, v in g["store"].items()} == {"event": 38, "cancel": 35, "state": 45}


def test_key60_is_the_reference_key_format():
    # first event key of the ring store and the (-1,-1) initial state key
    assert store.key60(0, 1, 0) == b"0000000000000000000\x000000000000000000001\x000000000000000000000\x00"
    assert store.key60(3, -1, -1) == b"0000000000000000003\x0000000000000000000-1\x0000000000000000000-1\x00"
    ks = [store.key60(0, -1, -1), store.key60(0, 1, 0), store.key60(0, 12, 2), store.key60(1, -1, -1)]
    assert ks == sorted(ks)          # '-1' sorts before '000' bytewise, as in the reference's LevelDB


def test_history_store_keys_and_file_roundtrip(tmp_path):
    """HistoryStore built from the init lines of the ring reproduces the reference's three key sets."""
    g = golden("ring_diff")
    recs = [O.parse_line(l) for l in g["init_lines"]]
    n = len(recs)
    rec = np.zeros(n, dtype=store.RAW_RECORD_DTYPE)
    sent = np.zeros(n, dtype=store.SENT_RECORD_DTYPE)
    succ = {(i, src, send) for i, src, send, dst, recv in recs}
    for k, (i, src, send, dst, recv) in enumerate(recs):
        rec[k]["key"] = (recv << 32) | i
        rec[k]["dst"] = dst
        rec[k]["src"] = 0xFFFFFFFF if src < 0 else src
        rec[k]["send_time"] = send
        sent[k]["dst"] = 0 if (i, dst, recv) in succ else store.NONE
    hs = store.HistoryStore(rec, sent, np.arange(10))
    assert hs.event_keys().tolist() == sorted(g["store"]["event"])
    assert hs.cancel_keys().tolist() == sorted(g["store"]["cancel"])
    assert hs.state_keys().tolist() == sorted(g["store"]["state"])
    t = hs.save_async(str(tmp_path), 999, 0)
    t.join()
    back = store.HistoryStore.load(str(tmp_path), 999, 0)
    assert back.rec.tobytes() == hs.rec.tobytes() and back.sent.tobytes() == hs.sent.tobytes()
    assert back.state_lps.tolist() == list(range(10))


def _store_from_lines(lines, state_lps):
    """HistoryStore of a run given by its result lines: the sent-log entry of an event is the line it produced (same id,
    source = this event's LP, send time = this event's receive time — the app convention)."""
    import scalesim_b200 as ssb
    parsed = [O.parse_line(l) for l in lines]                       # id, src, send, dst, recv
    produced = {(i, src, send): (dst, recv) for i, src, send, dst, recv in parsed if (src, send) != (dst, recv)}
    rec = np.zeros(len(parsed), dtype=ssb.RAW_RECORD_DTYPE)
    sent = np.zeros(len(parsed), dtype=ssb.SENT_RECORD_DTYPE)
    for k, (i, src, send, dst, recv) in enumerate(parsed):
        rec[k] = ((recv << 32) | i, dst, 0xFFFFFFFF if src < 0 else src, send, 0, 0, 0)
        out = produced.get((i, dst, recv))
        sent[k] = (out[0], i, out[1], 0) if out else (store.NONE, 0, 0, 0)
    return store.HistoryStore(rec, sent, np.asarray(state_lps, dtype=np.int64))


@pytest.mark.parametrize("case", [0, 1])
def test_store_has_no_records_for_lps_that_never_execute(case):
    """tests/golden/sparse_diff.json (make_golden_sparse.py, unmodified reference): networks where some junctions are
    never visited. The reference stores only LPs its scheduler ever dequeued (runner.hpp:573-583): an unvisited LP has
    no (lp, -1, -1) initial-state record. Event / cancel / state key sets derived by store.py == the reference's."""
    g = golden("sparse_diff")["cases"][case]
    hs = _store_from_lines(g["init_lines"], np.arange(g["n_nodes"]))
    assert hs.event_keys().tolist() == sorted(g["store"]["event"])
    assert hs.cancel_keys().tolist() == sorted(g["store"]["cancel"])
    assert hs.state_keys().tolist() == sorted(g["store"]["state"])
    assert hs.active_lps().tolist() == g["visited_lps"] and len(g["visited_lps"]) < g["n_nodes"]


def test_store_from_lines_reproduces_the_ring_store():
    """the same derivation on the ring (every LP visited): the reference's 38 / 35 / 45 records"""
    g = golden("ring_diff")
    hs = _store_from_lines(g["init_lines"], np.arange(10))
    assert (len(hs.event_keys()), len(hs.cancel_keys()), len(hs.state_keys())) == (38, 35, 45)
    assert hs.state_keys().tolist() == sorted(g["store"]["state"])
