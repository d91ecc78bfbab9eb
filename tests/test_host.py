"""CPU-side tests: the C-ABI library loads and exports every symbol include/scalesim_b200.h declares (no compute
without a GPU), the host mirrors of the two applications agree with the reference's own tests, and the Python /
oracle / golden digests are the same function."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib as O
import scalesim_b200 as ssb
from scalesim_b200 import phold, traffic
from fixtures import golden, golden_lines, ring_scenario, write_ring_files

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scalesim_b200.h")).read()
    names = set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", hdr))
    assert {"ssb_create", "ssb_run", "ssb_step", "ssb_load_events", "ssb_drain_committed", "ssb_comm_init"} <= names
    lib = ssb.load_library()
    for n in sorted(names):
        assert hasattr(lib, n), f"libscalesim_b200.so does not export {n}"
    assert lib.ssb_abi_version() == 2


def test_struct_layouts_match_header():
    assert ssb.EVENT_DTYPE.itemsize == 64 and ssb.RECORD_DTYPE.itemsize == 48
    assert ctypes.sizeof(ssb.capi._Config) == 88
    assert ctypes.sizeof(ssb.capi._Stats) == 14 * 8 + 8 + 3 * 8


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ssb.EngineError) as ei:
        ssb.Engine(ssb.Config(model=ssb.MODEL_PHOLD, n_lp=10, finish_time=100, bucket_ticks=10))
    assert ei.value.status == 2      # SSB_ERR_CUDA: there is no CPU fallback


def test_create_rejects_bad_config():
    for kw in (dict(n_lp=0), dict(finish_time=2**31), dict(bucket_ticks=0), dict(window_buckets=0), dict(model=99)):
        base = dict(model=ssb.MODEL_PHOLD, n_lp=10, finish_time=100, bucket_ticks=10)
        base.update(kw)
        with pytest.raises(ssb.EngineError) as ei:
            ssb.Engine(ssb.Config(**base))
        assert ei.value.status == 1


def test_phold_tables_match_oracle_and_pins(phold_tables_01):
    lat, rem = phold_tables_01
    assert lat[:10].tolist() == [14103, 61368, 24712, 113589, 272865, 73275, 3518, 75438, 772, 6918]
    assert rem[:10].tolist() == [0, 0, 0, 0, 0, 0, 1, 0, 1, 1]
    lat2 = np.empty_like(lat)
    rem2 = np.empty_like(rem)
    O.lib().two_phold_tables(1, 1.0, 0.1, lat.size, O._p(lat2), O._p(rem2))
    assert np.array_equal(lat, lat2) and np.array_equal(rem, rem2)
    assert int(lat.max()) < 2**31


def test_phold_init_events_identities(phold_tables_01):
    """reference test/large/phold/phold_test.cc:45-104: LATENCY_TABLE[id % size] is both send and receive time,
    source = destination = id % NUM_LP, and sharding by rank is a disjoint cover."""
    lat, _ = phold_tables_01
    num_lp, n = 100, 1600
    ev = phold.init_events(num_lp, n, lat)
    assert ev.size == n
    ids = ev["id"]
    assert np.array_equal(ids, np.arange(n))
    assert np.array_equal(ev["recv_time"], lat[ids % phold.RAND_TABLE_SIZE])
    assert np.array_equal(ev["send_time"], ev["recv_time"])
    assert np.array_equal(ev["dst"], ids % num_lp) and np.array_equal(ev["src"], ids % num_lp)
    assert not ev["p0"].any()
    shards = [phold.init_events(num_lp, n, lat, r, 4) for r in range(4)]
    assert sum(s.size for s in shards) == n
    assert np.array_equal(np.sort(np.concatenate([s["id"] for s in shards])), np.arange(n))
    for r, s in enumerate(shards):
        assert ((s["dst"] % 4) == r).all()          # owner = partition[dst] (phold.hpp:181-186)


def test_phold_partition_is_round_robin():
    """reference test/large/phold/phold_test.cc:25-43"""
    p = phold.partition(10, 4)
    assert p.tolist() == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]


def test_traffic_readers_roundtrip(tmp_path):
    """traffic_reader semantics (traffic_sim.hpp:345-451) on the ring inputs; reference test/small/io_test.cc
    pins the partition reader on the same ring files."""
    sc, g = ring_scenario()
    rd, trip, part = write_ring_files(str(tmp_path))
    sc2 = traffic.read_scenario(rd, trip, part)
    for k in ("state_ids", "road_start", "road_dst", "road_speed", "road_len", "trip_id", "trip_dep", "route_start",
              "route_nodes", "partition"):
        assert np.array_equal(getattr(sc, k), getattr(sc2, k)), k
    assert sc2.partition.tolist() == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]


def test_traffic_reader_quirks(tmp_path):
    """Quirks Q10/Q11: length 0 -> 1; state id = field 2 of the LAST road; trip field 2 ignored; ' 0' tolerated."""
    rd = tmp_path / "rd.csv"
    rd.write_text("R,0,7,1,10,0,1;R,1,5,2,0,300,2\n")
    trip = tmp_path / "trip.csv"
    trip.write_text("TP,42,999,17, 5,1,2\n")
    sc = traffic.read_scenario(str(rd), str(trip))
    assert sc.state_ids.tolist() == [5]
    assert sc.road_len.tolist() == [1, 300] and sc.road_speed.tolist() == [10, 0]
    assert sc.trip_id.tolist() == [42] and sc.trip_dep.tolist() == [17]
    assert sc.route_nodes.tolist() == [5, 1, 2]


def test_traffic_packing():
    sc, _ = ring_scenario()
    ids, words, table = traffic.pack_states(sc)
    assert words.shape == (10, traffic.STATE_WORDS) and table.size == 0
    assert words[3].tolist()[:5] == [3, 1, 4, 10, 100]
    # a junction with more out-roads than fit inline keeps them in the road table, in the order of the state's vectors
    big = traffic.Scenario(3, np.array([0, 1, 2]), np.array([0, 1, 5, 6]), np.array([1, 0, 2, 0, 2, 0]),
                           np.array([10, 20, 30, 40, 50, 60]), np.array([100, 200, 300, 400, 500, 600]),
                           np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(1, dtype=np.int64),
                           np.zeros(0, dtype=np.int64))
    ids, words, table = traffic.pack_states(big)
    assert words[1].tolist()[:3] == [1, 4, 0] and words[0].tolist()[:5] == [0, 1, 1, 10, 100]
    assert table.tolist() == [[0, 20, 200], [2, 30, 300], [0, 40, 400], [2, 50, 500]]
    ev = traffic.pack_events(sc)
    assert ev["src"].tolist() == [-1, -1, -1]
    assert ev["dst"].tolist() == [0, 1, 0]
    assert ev["p1"].tolist() == [11, 17, 10]
    assert ev["p0"].tolist() == [0, 11, 28]


def test_synthetic_grid_is_deterministic_and_valid():
    a = traffic.synthetic_grid(16, 16, 500, seed=1)
    b = traffic.synthetic_grid(16, 16, 500, seed=1)
    assert np.array_equal(a.route_nodes, b.route_nodes) and np.array_equal(a.road_len, b.road_len)
    assert a.road_speed.min() >= 10 and a.road_speed.max() <= 60
    assert a.road_len.min() >= 100 and a.road_len.max() <= 1000
    lens = np.diff(a.route_start)
    assert lens.min() >= 9 and lens.max() <= 65
    # every consecutive pair of route nodes is a road of the grid
    for t in range(0, 500, 37):
        r = a.route_nodes[a.route_start[t]:a.route_start[t + 1]]
        for u, v in zip(r[:-1], r[1:]):
            assert v in a.road_dst[a.road_start[u]:a.road_start[u + 1]]
    c = traffic.synthetic_grid(16, 16, 500, seed=2)
    assert not np.array_equal(a.route_nodes, c.route_nodes)


def test_stripe_partition():
    p = traffic.stripe_partition(4, 8, 4)
    assert p.size == 32 and p[:8].tolist() == [0] * 8 and p[-8:].tolist() == [3] * 8


def test_digest_functions_agree():
    lines = golden_lines("phold_100_1600")
    rec = np.zeros(len(lines), dtype=ssb.RECORD_DTYPE)
    for i, l in enumerate(lines):
        f = l.split(",")
        rec[i] = (int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), {"START": 1, "END": 2}.get(f[6], 0) if len(f) > 6 else 0)
    d_py = ssb.digest_records(rec)
    d_or = O.digest_records(rec.view(O.RECORD_DTYPE))
    assert d_py == d_or == golden("phold_100_1600")["digest"]
    assert ssb.records_to_lines(rec) == lines
    # order independence
    assert ssb.digest_records(rec[::-1]) == d_py
    # START / no-tag records (negative source) hash identically in both
    sc, g = ring_scenario()
    assert g["digest"][0] == 38


def test_what_if_reader_parses_the_reference_query_types(tmp_path):
    """traffic_reader::what_if_read (traffic_sim.hpp:453-599): AE / DE / RE lines; SC is refused."""
    from scalesim_b200 import traffic
    p = tmp_path / "wi.csv"
    p.write_text("AE,9990,0,20,0,1,2,3\nDE,4,165,0\nRE,2,2,94,3,135\n\nAE,7,0,5,8,9\n")
    wi = traffic.read_what_if(str(p))
    assert wi.trip_id.tolist() == [9990, 7] and wi.trip_dep.tolist() == [20, 5]
    assert wi.route_start.tolist() == [0, 4, 6] and wi.route_nodes.tolist() == [0, 1, 2, 3, 8, 9]
    assert wi.deletes.tolist() == [[4, 165, 0], [3, 135, 2]]          # (lp, time, event id)
    tid, dep, start, nodes = wi                                       # unpacks like the AE tuple
    assert nodes.size == 6
    # state change (traffic/ring/change_state.csv + the README's three-road example; length 0 -> 1 as in road_read)
    p.write_text("SC,1,5;R,1,1,2,10,200,1\nSC,10,555;R,0,10,1,10,100,2;R,1,10,1,10,0,1;R,3,10,3,10,200,100\n")
    wi = traffic.read_what_if(str(p))
    assert wi.state_changes == [(1, 5, [2], [10], [200]), (10, 555, [1, 1, 3], [10, 10, 10], [100, 1, 200])]


def test_history_store_keeps_the_model_data(tmp_path):
    from scalesim_b200 import store
    rec = np.zeros(3, dtype=store.RAW_RECORD_DTYPE)
    rec["key"] = [(5 << 32) | 1, (7 << 32) | 2, (9 << 32) | 1]
    rec["dst"] = [0, 1, 1]
    sent = np.zeros(3, dtype=store.SENT_RECORD_DTYPE)
    sent["dst"] = [1, store.NONE, 0]
    hs = store.HistoryStore(rec, sent, np.array([0, 1]), np.arange(10))
    hs.save(str(tmp_path), "sim", 3)
    back = store.HistoryStore.load(str(tmp_path), "sim", 3)
    assert back.pool.tolist() == list(range(10)) and back.state_lps.tolist() == [0, 1]
    assert back.event_keys().tolist() == [[0, 5, 1], [1, 7, 2], [1, 9, 1]]
    assert back.cancel_keys().tolist() == [[0, 5, 1], [1, 9, 1]]
    assert back.state_keys().tolist() == [[0, -1, -1], [0, 5, 1], [1, -1, -1], [1, 9, 1]]
    with pytest.raises(ValueError):
        (tmp_path / "sim" / "history3.ssb").write_bytes(b"garbage!" + bytes(64))
        store.HistoryStore.load(str(tmp_path), "sim", 3)


def test_compact_event_layout_roundtrip():
    ev = np.zeros(3, dtype=ssb.EVENT_DTYPE)
    ev["id"] = [1, 2, 3]; ev["src"] = [-1, 5, 6]; ev["dst"] = [4, 5, 6]
    ev["recv_time"] = [10, 20, 30]; ev["send_time"] = [10, 15, 25]; ev["p0"] = [0, 1, 2]; ev["flags"] = [0, 0, 1]
    raw = ssb.events_to_raw(ev)
    assert raw.dtype.itemsize == 32
    assert raw["key"].tolist() == [(10 << 32) | 1, (20 << 32) | 2, (30 << 32) | 3]
    assert raw["src"].tolist() == [0xFFFFFFFF, 5, 6] and raw["tag"].tolist() == [0, 0, 1]
    packed = np.zeros(2, dtype=ssb.PACKED_RECORD_DTYPE)
    packed["id"] = [1, 2]; packed["src"] = [0xFFFFFFFF, 3]; packed["recv_time"] = [7, 8]; packed["tag"] = [1, 2]
    rec = ssb.packed_to_records(packed)
    assert ssb.records_to_lines(rec) == ["RE,1,-1,0,0,7,START", "RE,2,3,0,0,8,END"]
