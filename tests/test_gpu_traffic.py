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
