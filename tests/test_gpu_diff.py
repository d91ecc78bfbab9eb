"""GPU: exact-differential simulation through the C ABI (ssb_drain_history / ssb_load_history) against the
reference's --diff_init / --diff_repeat runs and the repeat closure."""
import numpy as np
import pytest

import oracle_lib as O
import scalesim_b200 as ssb
from scalesim_b200 import store, traffic
from fixtures import golden, ring_scenario

pytestmark = pytest.mark.gpu


def _what_if(ae_lines, tmp_path):
    p = tmp_path / "ae.csv"
    p.write_text("".join(l + "\n" for l in ae_lines))
    return traffic.read_what_if(str(p))


def _grid(g):
    return traffic.synthetic_grid(**g["params"])


@pytest.mark.parametrize("case,bucket,window", [("ring", 8, 1), ("ring", 41, 4), ("grid", 8, 1), ("grid", 11, 1), ("grid", 16, 8)])
def test_diff_init_store_and_diff_repeat_equal_reference(case, bucket, window, tmp_path):
    g = golden(case + "_diff")
    if case == "ring":
        sc, _ = ring_scenario()
        sc.partition = None
    else:
        sc = _grid(g)
    # --diff_init: same output as a plain run, and the store holds the reference's logical records
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=bucket, window_buckets=window)
    assert sorted(ssb.records_to_lines(ssb.raw_to_records(rec))) == g["init_lines"]
    assert hs.event_keys().tolist() == sorted(g["store"]["event"])
    assert hs.cancel_keys().tolist() == sorted(g["store"]["cancel"])
    assert hs.state_keys().tolist() == sorted(g["store"]["state"])
    # the store survives the file round trip (written by a background thread)
    hs.save_async(str(tmp_path), 7, 0).join()
    hs = store.HistoryStore.load(str(tmp_path), 7, 0)
    # --diff_repeat with the AE queries
    eng = traffic.make_repeat_engine(sc, hs, _what_if(g["ae_lines"], tmp_path), bucket_ticks=bucket,
                                     window_buckets=window, keep_records=True)
    st2 = eng.run()
    lines = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    assert lines == g["repeat_union"]
    assert lines == O.repeat_closure(g["init_lines"], g["new_lines"])
    assert st2.rollbacks > 0 and st2.antis_sent > 0          # the history really was rolled back and cancelled
    assert st2.committed == len(lines) < len(g["init_lines"]) + len(g["new_lines"])


@pytest.mark.parametrize("case,bucket,window", [("ring", 8, 1), ("grid", 11, 1), ("grid", 16, 4)])
def test_delete_event_queries_equal_reference(case, bucket, window, tmp_path):
    """--diff_repeat with DE / RE (delete event) queries: traffic/ring/delete_ev.csv and three deletes on the grid."""
    g = golden(case + "_diff")
    if case == "ring":
        sc, _ = ring_scenario()
        sc.partition = None
    else:
        sc = _grid(g)
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=bucket, window_buckets=window)
    wi = _what_if(g["de_lines"], tmp_path)
    assert wi.deletes.shape == (len(g["de_lines"]), 3) and wi.trip_id.size == 0
    eng = traffic.make_repeat_engine(sc, hs, wi, bucket_ticks=bucket, window_buckets=window, keep_records=True)
    st2 = eng.run()
    lines = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    assert lines == g["delete_union"]
    assert st2.annihilated >= len(g["de_lines"])


def test_repeat_is_exact_on_a_larger_grid(tmp_path):
    """Config-5 shape at a reduced size: init + repeat(AE) must agree with one full run over trips + added trips
    (vehicles do not interact, so full = init + new; repeat = new + re-executed history = the closure)."""
    sc = traffic.synthetic_grid(64, 64, 40000, seed=2)
    extra = traffic.synthetic_grid(64, 64, 500, seed=9, max_departure=6000)
    extra_ids = 10_000_000 + np.arange(extra.trip_id.size)
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=11, window_buckets=1)
    init_lines = sorted(ssb.records_to_lines(ssb.raw_to_records(rec)))
    # full run: both trip sets at once
    full = traffic.Scenario(sc.n_lp, sc.state_ids, sc.road_start, sc.road_dst, sc.road_speed, sc.road_len,
                            np.concatenate([sc.trip_id, extra_ids]), np.concatenate([sc.trip_dep, extra.trip_dep]),
                            np.concatenate([sc.route_start, sc.route_start[-1] + extra.route_start[1:]]),
                            np.concatenate([sc.route_nodes, extra.route_nodes]))
    eng = traffic.make_engine(full, bucket_ticks=11, keep_records=True)
    eng.run()
    full_lines = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    new_lines = sorted(set(full_lines) - set(init_lines))
    assert sorted(set(full_lines) - set(new_lines)) == init_lines           # added trips change nothing else
    eng = traffic.make_repeat_engine(sc, hs, (extra_ids, extra.trip_dep, extra.route_start, extra.route_nodes),
                                     bucket_ticks=11, window_buckets=1, keep_records=True)
    st2 = eng.run()
    rep = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    assert sorted(set(rep) - set(init_lines)) == new_lines
    assert set(rep) <= set(full_lines) and len(set(rep)) == len(rep)
    assert rep == O.repeat_closure(init_lines, new_lines)
    assert st2.committed < len(full_lines)                                   # less work than the full run


@pytest.mark.parametrize("window", [1, 3])
def test_phold_diff_repeat_equals_reference(window):
    """PHOLD 100/1600 --diff_init, then --diff_repeat with 3 what-if events (phold.hpp:232-246: id NUM_INIT_MSG + i
    at LP i, time 1)."""
    from scalesim_b200 import phold
    g = golden("phold_diff")
    n_lp, n_init, ratio, lam = g["args"]
    tables = phold.build_tables(ratio, lam)
    cfg = dict(tables=tables, window_buckets=window, max_committed=20000)
    eng = phold.make_engine(n_lp, n_init, ratio, lam, flags=ssb.FLAG_HISTORY, **cfg)
    eng.run()
    rec, sent = eng.drain_history()
    eng.close()
    assert sorted(ssb.records_to_lines(ssb.raw_to_records(rec))) == g["init_lines"]
    eng = phold.make_engine(n_lp, n_init, ratio, lam, load=False, max_events=60000, **cfg)
    eng.load_history(rec, sent)
    ev = np.zeros(g["what_ifs"], dtype=ssb.EVENT_DTYPE)
    ev["id"] = n_init + np.arange(g["what_ifs"])
    ev["src"] = ev["dst"] = np.arange(g["what_ifs"])
    ev["recv_time"] = ev["send_time"] = 1
    eng.load_events(ev)
    st = eng.run()
    lines = sorted(ssb.records_to_lines(eng.drain_committed()))
    eng.close()
    assert lines == g["repeat_union"] == g["repeat_core"]
    assert st.rollbacks > 0


def test_history_needs_the_flag():
    sc, _ = ring_scenario()
    sc.partition = None
    eng = traffic.make_engine(sc, keep_records=True)
    eng.run()
    with pytest.raises(ssb.EngineError):
        eng.drain_history()
    eng.close()


def test_streaming_history_store_equals_the_drained_one(tmp_path):
    """ssb_run_stream_history: the history leaves through the pinned staging buffers while the loop runs and a writer
    thread appends it to the store file — same records as ssb_drain_history after the run."""
    sc = traffic.synthetic_grid(48, 48, 30000, seed=4)
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=11)
    st2, path = traffic.run_diff_init_streaming(sc, str(tmp_path), 5, bucket_ticks=11)
    hs2 = store.HistoryStore.load(str(tmp_path), 5, 0)
    assert st2.committed == st.committed == hs2.rec.size
    a = np.sort(np.stack([hs.rec["key"], hs.rec["dst"].astype(np.uint64), hs.sent["dst"].astype(np.uint64)], axis=1), axis=0)
    k1 = np.lexsort((hs.sent["recv_time"], hs.sent["dst"], hs.rec["dst"], hs.rec["key"]))
    k2 = np.lexsort((hs2.sent["recv_time"], hs2.sent["dst"], hs2.rec["dst"], hs2.rec["key"]))
    assert np.array_equal(hs.rec[k1], hs2.rec[k2]) and np.array_equal(hs.sent[k1], hs2.sent[k2])
    assert np.array_equal(hs2.pool, sc.route_nodes) and hs2.state_lps.size == sc.state_ids.size
    # a sink that fails stops the run loudly
    eng = traffic.make_engine(sc, bucket_ticks=11, history=True)
    with pytest.raises(ssb.EngineError):
        eng.run_stream_history(lambda rec, sent: 7)
    eng.close()


def ring_changed_network_lines():
    """one plain run (oracle closure) of the ring example on the network of change_state.csv: LP 1's road = 200 m"""
    sc, _ = ring_scenario()
    changed = traffic.Scenario(sc.n_lp, sc.state_ids, sc.road_start, sc.road_dst, sc.road_speed, sc.road_len.copy(),
                               sc.trip_id, sc.trip_dep, sc.route_start, sc.route_nodes)
    changed.road_len[changed.road_start[1]] = 200
    o = O.Sim.traffic_scenario(changed)
    o.run_closure(2)
    return sorted(O.records_to_lines(o.committed()))


def test_state_change_query_on_the_ring(tmp_path):
    """SC what-if (traffic/ring/change_state.csv: LP 1's road gets 200 m from t = 5, before any event reaches LP 1).
    Exactness (the definition of exact-differential simulation): init output minus what the change supersedes, plus
    the new lines of the repeat run, must equal ONE run on the changed network.
    The reference's own SC output is not a parity target as a whole: it drops every anti-message that reaches an LP
    before that LP's lazy history load (logical_process.hpp:132-138, queue.hpp:87-90), so it keeps the stale
    trajectories (30 / 31 stale lines, 69 lines at 2 worker threads, 70 at 3 / 4 / 8) and even re-executes a stale
    arrival at LP 1 with the new state (a 6-line trajectory no run on the changed network has). What IS checked against
    it: this engine's output is contained in every reference variant, and its new lines are exactly the reference's
    new lines that are right (= also lines of the run on the changed network)."""
    g = golden("ring_diff")
    sc, _ = ring_scenario()
    sc.partition = None
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=8)
    init_lines = sorted(ssb.records_to_lines(ssb.raw_to_records(rec)))
    assert init_lines == g["init_lines"]
    wi = _what_if(g["sc_lines"], tmp_path)
    assert len(wi.state_changes) == 1 and wi.state_changes[0][:2] == (1, 5) and wi.state_changes[0][4] == [200]
    full = ring_changed_network_lines()
    superseded = set(init_lines) - set(full)
    for bucket, window in ((8, 1), (41, 3)):
        eng = traffic.make_repeat_engine(sc, hs, wi, bucket_ticks=bucket, window_buckets=window, keep_records=True)
        st2 = eng.run()
        lines = sorted(ssb.records_to_lines(eng.drain_committed()))
        eng.close()
        new = sorted(l for l in lines if l not in set(init_lines))
        assert sorted((set(init_lines) - superseded) | set(new)) == full                  # exact
        assert new == sorted(set(g["sc_new_lines"]) & set(full))                          # the reference's right new lines
        assert len(set(g["sc_new_lines"]) - set(full)) == 6                               # ... its artefact trajectory
        for v in g["sc_variants"]:
            assert set(lines) <= set(v["lines"])
        assert st2.rollbacks > 0 and st2.antis_sent >= len(superseded)


def test_state_change_query_with_more_roads_than_fit_inline(tmp_path):
    """An SC query whose new road set has more out-roads than the state record holds inline: the roads go to the road
    table behind the scenario's own entries (make_repeat_engine: extra_roads). Four dead-end roads in front of the one
    the routes use (std::find takes the first road that leads to the next node, traffic_sim.hpp:299-306): the result
    is the run on the network where LP 1's road is 200 m long."""
    sc, _ = ring_scenario()
    sc.partition = None
    st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=8)
    init_lines = sorted(ssb.records_to_lines(ssb.raw_to_records(rec)))
    wi = _what_if(["SC,1,5;R,1,1,5,10,50,1;R,1,1,6,10,60,1;R,1,1,7,10,70,1;R,1,1,8,10,80,1;R,1,1,2,10,200,1"], tmp_path)
    assert len(wi.state_changes[0][2]) == 5 > traffic.INLINE_ROADS
    full = ring_changed_network_lines()
    superseded = set(init_lines) - set(full)
    for bucket, window in ((8, 1), (41, 3)):
        eng = traffic.make_repeat_engine(sc, hs, wi, bucket_ticks=bucket, window_buckets=window, keep_records=True)
        eng.run()
        lines = sorted(ssb.records_to_lines(eng.drain_committed()))
        eng.close()
        new = set(l for l in lines if l not in set(init_lines))
        assert sorted((set(init_lines) - superseded) | new) == full


@pytest.mark.parametrize("case", [0, 1])
def test_store_of_a_network_with_unvisited_junctions_equals_reference(case):
    """--diff_init on the networks of tests/golden/sparse_diff.json: the engine's history gives the reference's three
    key sets — in particular no initial-state record for a junction no event ever reached (runner.hpp:573-583)."""
    from fixtures import sparse_network
    g = golden("sparse_diff")["cases"][case]
    sc = sparse_network(g["n_nodes"], g["n_trips"], g["seed"])
    sc.partition = None
    for bucket, window in ((8, 1), (50, 3)):
        st, rec, hs = traffic.run_diff_init(sc, bucket_ticks=bucket, window_buckets=window)
        assert sorted(ssb.records_to_lines(ssb.raw_to_records(rec))) == g["init_lines"]
        assert hs.event_keys().tolist() == sorted(g["store"]["event"])
        assert hs.cancel_keys().tolist() == sorted(g["store"]["cancel"])
        assert hs.state_keys().tolist() == sorted(g["store"]["state"])
