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
