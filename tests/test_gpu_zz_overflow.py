"""Capacity overflows on the device must end a run with SSB_ERR_CAPACITY — never with a wild access by the rounds that
are already enqueued when the overflow happens (found by tools/emu/fuzz.py under AddressSanitizer: a bucket whose fill
count points past its chunk list; an irregular LP's segment that does not fit the entry table). The file sorts last on
purpose: if one of these ever faults on a GPU, every other parity test has run before."""
import numpy as np
import pytest

import oracle_lib as O
import scalesim_b200 as ssb
from scalesim_b200 import phold, traffic
from test_gpu_traffic import _random_network

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", ["graph", "stream"])
def test_event_pool_overflow_is_a_clean_capacity_error(path, monkeypatch):
    """A window far wider than a hop (bucket 4000 s, two buckets) makes every round's outputs stragglers of the next:
    the re-sent copies pile up in the two uncommitted buckets until the event pool is exhausted. That must surface as
    SSB_ERR_CAPACITY from ssb_run — not as a wild access of the rounds that are already enqueued when the overflow
    happens (a bucket's fill count then points past its chunk list) — and the engine must be usable after a reset."""
    if path == "stream":
        monkeypatch.setenv("SSB_NO_GRAPH", "1")
    sc = _random_network(40, 1, 2000, seed=567568399)
    eng = traffic.make_engine(sc, bucket_ticks=4000, window_buckets=2)
    with pytest.raises(ssb.EngineError) as ei:
        eng.run()
    assert ei.value.status == 4 and "event pool exhausted" in ei.value.message
    # same engine, a load it can hold: the first 60 trips
    small = traffic.Scenario(sc.n_lp, sc.state_ids, sc.road_start, sc.road_dst, sc.road_speed, sc.road_len, sc.trip_id[:60],
                             sc.trip_dep[:60], sc.route_start[:61], sc.route_nodes[:sc.route_start[60]])
    s = O.Sim.traffic_scenario(small)
    s.run_closure(2)
    eng.reset()
    eng.load_events(traffic.pack_events(small))
    st = eng.run()
    assert eng.digest() == s.digest() and st.rollbacks > 0
    eng.close()


def test_event_pool_too_small_for_the_load_fails_at_load():
    sc = traffic.synthetic_grid(16, 16, 100000, seed=4)
    with pytest.raises(ssb.EngineError) as ei:
        traffic.make_engine(sc, bucket_ticks=8, max_events=30000)
    assert ei.value.status == 4


@pytest.mark.parametrize("path", ["graph", "stream"])
def test_rollback_segment_overflow_is_a_clean_capacity_error(path, monkeypatch):
    """One LP, 12 000 events with short hops (lambda 30), a window of 200 000 ticks: the LP's rollback segment (arrivals + undone events) outgrows
    the entry table (max_batch + auxiliary entries). SSB_ERR_CAPACITY, and the engine works again after a reset with a
    window it can hold."""
    if path == "stream":
        monkeypatch.setenv("SSB_NO_GRAPH", "1")
    tables = phold.build_tables(0.5, 30.0)
    eng = phold.make_engine(1, 12000, 0.5, 30.0, tables=tables, bucket_ticks=200000, window_buckets=1,
                            max_events=12000 * 3 + (1 << 16), max_batch=12000 + (1 << 14), load=False)
    ev = phold.init_events(1, 12000, tables[0])
    eng.load_events(ev)
    with pytest.raises(ssb.EngineError) as ei:
        eng.run()
    assert ei.value.status == 4 and "raise max_batch" in ei.value.message
    eng.reset()
    eng.load_events(ev[:300])
    st = eng.run()
    s = O.Sim.phold(1, 300, 0.5, 30.0)
    s.run_closure(2)
    assert eng.digest() == s.digest() and st.rollbacks > 0
    eng.close()
