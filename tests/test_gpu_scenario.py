"""Input at scale on the device: an engine started from a scenario image (ssb_scenario_load — one
mapping, a handful of host-to-device copies) commits exactly what an engine fed through the per-array calls commits,
which the other parity tests pin on the reference."""
import os

import numpy as np
import pytest

import oracle_lib as O
import scalesim_b200 as ssb
from scalesim_b200 import phold, scenario, traffic
from fixtures import golden, ring_scenario, write_ring_files

pytestmark = pytest.mark.gpu


def _bare_engine(model, n_lp, finish_time, bucket_ticks, n_events, *, window=1, max_committed=0, rank=0, world=1):
    return ssb.Engine(ssb.Config(model=model, n_lp=n_lp, finish_time=finish_time, bucket_ticks=bucket_ticks,
                                 window_buckets=window, max_events=int(n_events * 1.6) + (1 << 16),
                                 max_batch=max(1 << 14, n_events // 2 + (1 << 14)), max_committed=max_committed,
                                 rank=rank, world=world))


@pytest.mark.parametrize("window", [1, 3])
def test_phold_from_an_image_equals_reference(tmp_path, phold_tables_01, window):
    """PHOLD 100 LPs / 1600 events from a saved and re-opened image: the reference's 7 517 lines."""
    g = golden("phold_100_1600")
    p = str(tmp_path / "phold.ssbscn")
    scenario.phold_image(100, 1600, 0.1, 1.0, tables=phold_tables_01).save(p)
    with scenario.Image.open(p) as im:
        eng = _bare_engine(ssb.MODEL_PHOLD, im.n_lp, im.finish_time, phold.LOOK_AHEAD, im.n_events, window=window, max_committed=20000)
        eng.load_scenario(im)
    st = eng.run()                       # the mapping is gone: everything was copied by the load
    assert st.committed == g["digest"][0] == 7517
    assert eng.digest() == g["digest"]
    assert (st.rollbacks > 0) == (window > 1)
    eng.close()


def test_ring_from_csv_through_a_compiled_image_equals_reference(tmp_path):
    """traffic/ring CSV files -> ssb_traffic_compile -> image -> engine: the reference's 38 lines."""
    g = golden("ring")
    rd, trip, part = write_ring_files(str(tmp_path))
    img = str(tmp_path / "ring.ssbscn")
    scenario.compile_traffic_csv(rd, trip, img, partition_file=part)
    with scenario.Image.open(img) as im:
        assert im.n_lp == 10 and im.n_events == 3
        eng = _bare_engine(ssb.MODEL_TRAFFIC, im.n_lp, im.finish_time, 8, im.n_events, max_committed=1024)
        eng.load_scenario(im)            # world = 1: the 4-way partition file reduces to one partition
    st = eng.run()
    assert st.committed == 38
    assert sorted(ssb.records_to_lines(eng.drain_committed())) == g["lines"]
    assert eng.digest() == g["digest"]
    eng.close()


def test_grid_with_junctions_of_any_degree_from_an_image_equals_oracle(tmp_path):
    """A random network with up to 5 out-roads per junction (road table in model data slot 1), written as CSV, parsed
    and packed natively: digest = the oracle's closure of the same scenario."""
    from test_gpu_traffic import _random_network
    sc = _random_network(300, 5, 2000, seed=11)
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(2)
    rd, trip, img = (str(tmp_path / n) for n in ("rd.csv", "trip.csv", "net.ssbscn"))
    scenario.write_traffic_csv(sc, rd, trip)
    scenario.compile_traffic_csv(rd, trip, img, threads=3)
    with scenario.Image.open(img) as im:
        assert im.slot(1) is not None
        eng = _bare_engine(ssb.MODEL_TRAFFIC, im.n_lp, im.finish_time, 8, im.n_events, window=2)
        eng.load_scenario(im)
    eng.run()
    assert eng.digest() == s.digest()
    eng.close()


def test_image_load_keeps_only_this_ranks_share():
    """world = 2 (no exchange set up; nothing is run): each engine takes the states of its own LPs and only events for
    them — runner::init's split by partition[destination] (runner.hpp:131-147)."""
    sc = traffic.synthetic_grid(8, 8, 400, seed=2)
    sc.partition = traffic.stripe_partition(8, 8, 4)                 # 4 stripes on 2 ranks: owner = stripe % 2
    im = scenario.traffic_image(sc)
    ids, words, _ = traffic.pack_states(sc)
    ev = traffic.pack_events(sc)
    total = 0
    for rank in (0, 1):
        eng = _bare_engine(ssb.MODEL_TRAFFIC, sc.n_lp, traffic.FINISH_TIME, 8, 400, rank=rank, world=2)
        eng.load_scenario(im)
        mine = sc.partition[ids] % 2 == rank
        assert np.array_equal(eng.read_states(ids[mine]), words[mine])
        with pytest.raises(ssb.EngineError):
            eng.read_states(ids[~mine][:1])
        st = eng.stats()
        assert st.remote_sent == 0
        n_mine = int((sc.partition[ev["dst"]] % 2 == rank).sum())
        assert st.pool_high_water >= n_mine > 0
        total += n_mine
        eng.close()
    assert total == 400


def test_image_for_another_model_or_size_is_refused(phold_tables_01):
    im = scenario.phold_image(64, 256, 0.1, 1.0, tables=phold_tables_01)
    eng = _bare_engine(ssb.MODEL_TRAFFIC, 64, 10800, 8, 256)
    with pytest.raises(ssb.EngineError, match="model"):
        eng.load_scenario(im)
    eng.close()
    eng = _bare_engine(ssb.MODEL_PHOLD, 65, phold.FINISH_TIME, phold.LOOK_AHEAD, 256)
    with pytest.raises(ssb.EngineError, match="64 LPs"):
        eng.load_scenario(im)
    eng.close()


def test_ingest_measure_small():
    """What bench.py reports under the `ingest` key, at a small size: CSV -> native readers -> image -> engine commits
    what the array path commits."""
    res = scenario.ingest_measure(32, 32, 3000, bucket_ticks=11)
    assert res["read_equals_generator"] and res["digest_equal"]
    s = O.Sim.traffic_scenario(traffic.synthetic_grid(32, 32, 3000, seed=1))
    s.run_closure(2)
    assert res["committed"] == s.digest()[0]
    for k in ("write_csv_s", "native_read_s", "compile_to_image_s", "open_image_s", "load_image_s", "array_path_pack_and_load_s"):
        assert res[k] >= 0
