"""GPU parity tests for TrafficSim (through the C ABI) vs the reference's ring output and the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import scalesim_b200 as ssb
from scalesim_b200 import traffic
from fixtures import ring_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bucket,window", [(8, 1), (41, 1), (8, 16), (100, 200), (1, 1)])
def test_ring_equals_reference(bucket, window):
    sc, g = ring_scenario()
    sc.partition = None
    eng = traffic.make_engine(sc, bucket_ticks=bucket, window_buckets=window, keep_records=True)
    st = eng.run()
    rec = eng.drain_committed()
    assert st.committed == 38
    assert sorted(ssb.records_to_lines(rec)) == g["lines"]
    assert eng.digest() == g["digest"]
    eng.close()


@pytest.mark.parametrize("bucket,window", [(8, 1), (8, 8), (64, 4)])
def test_synthetic_grid_equals_oracle(bucket, window):
    sc = traffic.synthetic_grid(32, 32, 20000, seed=1)
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(4)
    want = s.digest()
    s2 = O.Sim.traffic_scenario(sc)
    s2.run_timewarp(4, 100, 5)
    assert s2.digest() == want
    eng = traffic.make_engine(sc, bucket_ticks=bucket, window_buckets=window, keep_records=True)
    st = eng.run()
    assert eng.digest() == want
    rec = eng.drain_committed()
    assert sorted(ssb.records_to_lines(rec)) == sorted(O.records_to_lines(s.committed()))
    if window * bucket > 11:
        assert st.rollbacks > 0
    eng.close()


def test_grid_trips_that_outlive_finish_time():
    """Events at or beyond finish_time are never committed (R8), even when executed."""
    sc = traffic.synthetic_grid(16, 16, 3000, seed=3, max_departure=10700)
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(2)
    eng = traffic.make_engine(sc, bucket_ticks=16, window_buckets=2)
    st = eng.run()
    assert eng.digest() == s.digest()
    assert st.beyond_finish > 0
    eng.close()


def _random_network(n_nodes, max_degree, n_trips, seed):
    """random directed road network with 1..max_degree out-roads per junction (duplicate destinations allowed: the
    handler takes the first match, traffic_sim.hpp:299-306) and random walks over it as trips"""
    rng = np.random.default_rng(seed)
    deg = rng.integers(1, max_degree + 1, n_nodes)
    road_start = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(deg, out=road_start[1:])
    road_dst = rng.integers(0, n_nodes, road_start[-1])
    road_speed = rng.integers(0, 61, road_start[-1])          # 0 = the reference's "speed unknown" branch (:318-319)
    road_len = rng.integers(1, 1200, road_start[-1])
    trip_id, trip_dep, route_start, nodes = [], [], [0], []
    for t in range(n_trips):
        cur = int(rng.integers(0, n_nodes))
        route = [cur]
        for _ in range(int(rng.integers(1, 12))):
            cur = int(road_dst[road_start[cur] + rng.integers(0, deg[cur])])
            route.append(cur)
        if rng.random() < 0.1:
            route.append(int(rng.integers(0, n_nodes)))          # most likely no road leads there: the trip ends early
        trip_id.append(t); trip_dep.append(int(rng.integers(0, 3000)))
        nodes.extend(route); route_start.append(len(nodes))
    a = lambda x: np.array(x, dtype=np.int64)
    return traffic.Scenario(n_nodes, np.arange(n_nodes, dtype=np.int64), road_start, a(road_dst), a(road_speed), a(road_len),
                            a(trip_id), a(trip_dep), a(route_start), a(nodes))


@pytest.mark.parametrize("bucket,window", [(5, 1), (40, 4)])
def test_any_junction_degree_equals_oracle(bucket, window):
    """Junctions with more out-roads than fit the inline state words (road table, model data slot 1): the reference's
    State holds unbounded vectors (traffic_sim.hpp:99-142)."""
    sc = _random_network(300, 9, 4000, seed=5)
    assert np.diff(sc.road_start).max() > traffic.INLINE_ROADS
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(4)
    eng = traffic.make_engine(sc, bucket_ticks=bucket, window_buckets=window, keep_records=True)
    st = eng.run()
    assert eng.digest() == s.digest()
    rec = eng.drain_committed()
    assert sorted(ssb.records_to_lines(rec)) == sorted(O.records_to_lines(s.committed()))
    eng.close()


def test_traffic_config3_full_size_count_and_digest():
    """BASELINE.json config 3 at full size: 1024 x 1024 torus grid, 10M trips -> 367 737 881 committed events, digest =
    the oracle's closure (every trip advanced hop by hop on the CPU). Also the size-independent property of a window
    within the lookahead: every event is processed exactly once, no rollback."""
    sc = traffic.synthetic_grid(1024, 1024, 10_000_000, seed=1)
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(8, keep_records=False)
    want = s.digest()
    s.close()
    n = sc.trip_id.size
    eng = traffic.make_engine(sc, bucket_ticks=11, window_buckets=1, max_events=2 * n + (1 << 20), max_batch=n + (1 << 16))
    del sc
    st = eng.run()
    assert st.committed == 367_737_881 == want[0]
    assert st.processed == st.committed and st.rollbacks == 0
    assert eng.digest() == want
    eng.close()

