"""Input at scale, host side — no GPU: the native TrafficSim readers (ssb_traffic_read) against the
per-line restatement of the reference's readers (scalesim_b200/traffic.py, traffic_sim.hpp:345-451) and against the
unmodified reference binary where it is built; the scenario image (ssb_scenario_*) round trip and its loud failures."""
import os
import subprocess

import numpy as np
import pytest

from fixtures import golden, ring_scenario, write_ring_files

from scalesim_b200 import EngineError, MODEL_PHOLD, MODEL_TRAFFIC, RAW_RECORD_DTYPE, events_to_raw, phold, scenario, traffic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("state_ids", "road_start", "road_dst", "road_speed", "road_len", "trip_id", "trip_dep", "route_start", "route_nodes")


def _python_tables(rd, trip, part=None):
    sid, rs, rdst, sp, ln = traffic.read_roads(rd)
    tid, dep, ts, nodes = traffic.read_trips(trip)
    out = dict(state_ids=sid, road_start=rs, road_dst=rdst, road_speed=sp, road_len=ln, trip_id=tid, trip_dep=dep,
               route_start=ts, route_nodes=nodes)
    out["part"] = traffic.read_partition(part) if part else None
    return out


def _same(a, b):
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert (a["part"] is None) == (b["part"] is None)
    if a["part"] is not None:
        assert np.array_equal(a["part"], b["part"])


def test_native_readers_equal_the_line_readers_on_the_ring(tmp_path):
    rd, trip, part = write_ring_files(str(tmp_path))
    sc, g = ring_scenario()
    for threads in (1, 3, 0):
        nat = scenario.read_traffic_csv(rd, trip, part, threads=threads)
        _same(nat, _python_tables(rd, trip, part))
        for k in KEYS:
            assert np.array_equal(nat[k], getattr(sc, k)), k            # the array fixture the goldens were made from
        assert np.array_equal(nat["part"], sc.partition)


def test_native_readers_edge_cases(tmp_path):
    """atol semantics (blanks, signs, trailing text, empty fields), '\\r\\n' line ends, a last line without '\\n',
    short lines skipped, junctions of degree 1 and 3, a zero length that becomes 1."""
    rd, trip, part = (str(tmp_path / n) for n in ("rd.csv", "trip.csv", "part"))
    with open(rd, "w", newline="") as f:
        f.write("R,0,0,1,40,500,1;R,1,0,2, 30km,0,2\r\n"          # two roads, '30km' -> 30, length 0 -> 1, CR stays in the lanes field
                "R,2,1,2,+50,-7,1\n"                                 # signs
                "R,3,2,0,,100,1;R,4,2,1,60,200,1;R,5,2,2,10,300,1;short,road\n"   # empty speed -> 0, a short road is skipped
                "R,6,3,0,20,400,1")                                  # no newline at the end
    with open(trip, "w", newline="") as f:
        f.write("TP,7,0,100,0,1,2\n"
                "too,short,line,1\n"                                 # 4 fields: skipped
                "TP, 8 ,0,\t200x,2\r\n"                              # blanks, trailing text, CRLF
                "TP,9,0,300,\n"                                      # empty route field -> node 0
                "TP,10,0,400,3,0")
    with open(part, "w") as f:
        f.write("0\n1\n 2\nx\n3")
    for threads in (1, 2, 5):
        nat = scenario.read_traffic_csv(rd, trip, part, threads=threads)
        _same(nat, _python_tables(rd, trip, part))
    assert nat["state_ids"].tolist() == [0, 1, 2, 3]
    assert nat["road_start"].tolist() == [0, 2, 3, 6, 7]
    assert nat["road_speed"].tolist() == [40, 30, 50, 0, 60, 10, 20]
    assert nat["road_len"].tolist() == [500, 1, -7, 100, 200, 300, 400]
    assert nat["trip_id"].tolist() == [7, 8, 9, 10] and nat["trip_dep"].tolist() == [100, 200, 300, 400]
    assert nat["route_start"].tolist() == [0, 3, 4, 5, 7] and nat["route_nodes"].tolist() == [0, 1, 2, 2, 0, 3, 0]
    assert nat["part"].tolist() == [0, 1, 2, 0, 3]


def test_native_readers_on_a_generated_grid_many_pieces(tmp_path):
    """A 48x48 torus with 20 000 trips written in the reference's format (ssb_traffic_write_csv), read back with more
    threads than the file has MiB (pieces are cut at line ends wherever they fall)."""
    sc = traffic.synthetic_grid(48, 48, 20_000, seed=5)
    sc.partition = traffic.stripe_partition(48, 48, 4)
    rd, trip, part = (str(tmp_path / n) for n in ("rd.csv", "trip.csv", "part"))
    scenario.write_traffic_csv(sc, rd, trip, part)
    py = _python_tables(rd, trip, part)
    for k in KEYS:
        assert np.array_equal(py[k], getattr(sc, k)), k                # writer -> line reader = generator
    for threads in (1, 2, 7):
        _same(scenario.read_traffic_csv(rd, trip, part, threads=threads), py)
    assert os.path.getsize(trip) > (1 << 20)                           # several pieces really are used


def test_missing_file_fails_loudly(tmp_path):
    with pytest.raises(EngineError, match="cannot open trip file"):
        scenario.read_traffic_csv(None, str(tmp_path / "nope.csv"))
    with pytest.raises(EngineError, match="cannot open scenario image"):
        scenario.Image.open(str(tmp_path / "nope.ssbscn"))


def test_compiled_image_equals_the_python_packing(tmp_path):
    """ssb_traffic_compile (parse + pack + save, native) == traffic.pack_states / pack_events of the same scenario;
    a junction of degree 3 goes through the road table."""
    sc = traffic.synthetic_grid(16, 16, 500, seed=3)
    # give node 5 a third road: its roads move to the road table
    k = int(sc.road_start[6])
    sc.road_dst = np.insert(sc.road_dst, k, 7)
    sc.road_speed = np.insert(sc.road_speed, k, 33)
    sc.road_len = np.insert(sc.road_len, k, 444)
    sc.road_start = sc.road_start.copy()
    sc.road_start[6:] += 1
    sc.partition = traffic.stripe_partition(16, 16, 2)
    rd, trip, part, img = (str(tmp_path / n) for n in ("rd.csv", "trip.csv", "part", "grid.ssbscn"))
    scenario.write_traffic_csv(sc, rd, trip, part)
    scenario.compile_traffic_csv(rd, trip, img, partition_file=part, finish_time=10800, threads=2)
    ids, words, table = traffic.pack_states(sc)
    with scenario.Image.open(img) as im:
        assert (im.model, im.n_lp, im.finish_time) == (MODEL_TRAFFIC, 256, 10800)
        assert np.array_equal(im.partition, sc.partition)
        assert np.array_equal(im.state_lps, ids) and np.array_equal(im.states, words)
        assert np.array_equal(im.slot(0), sc.route_nodes) and np.array_equal(im.slot(1), table.reshape(-1))
        assert table.shape == (3, 3) and im.slot(2) is None
        assert np.array_equal(im.events, events_to_raw(traffic.pack_events(sc)))
    # the Python-side image of the same scenario is the same file, byte for byte
    img2 = str(tmp_path / "grid2.ssbscn")
    scenario.traffic_image(sc).save(img2)
    assert open(img, "rb").read() == open(img2, "rb").read()


def test_phold_image_round_trip(tmp_path):
    lat = np.arange(1000, dtype=np.int64) * 37 % 9001 + 1
    rem = (np.arange(1000) % 10 == 0).astype(np.int32)
    im = scenario.phold_image(64, 1024, 0.1, 1.0, tables=(lat, rem))
    p = str(tmp_path / "phold.ssbscn")
    im.save(p)
    assert os.path.getsize(p) % 4096 == 0
    with scenario.Image.open(p) as back:
        assert (back.model, back.n_lp, back.finish_time, back.n_events, back.n_states) == (MODEL_PHOLD, 64, phold.FINISH_TIME, 1024, 64)
        assert back.partition is None
        assert np.array_equal(back.slot(0), lat) and np.array_equal(back.slot(1, np.int32), rem)
        assert back.slot(2).tolist() == [phold.LOOK_AHEAD, 64]
        assert np.array_equal(back.events, events_to_raw(phold.init_events(64, 1024, lat)))
        assert back.events.dtype == RAW_RECORD_DTYPE and back.events.ctypes.data % 4096 == 0      # page-aligned section


def test_damaged_image_is_rejected(tmp_path):
    p = str(tmp_path / "a.ssbscn")
    scenario.traffic_image(ring_scenario()[0]).save(p)
    raw = open(p, "rb").read()
    cases = {
        "magic": b"SSBSCN99" + raw[8:],
        "size in the header": raw[:-4096],                                    # truncated
        "outside the file": raw[:8 + 6 * 8] + (10 ** 12).to_bytes(8, "little") + raw[8 + 7 * 8:],      # n_events
        "too short": raw[:100],
    }
    for why, data in cases.items():
        q = str(tmp_path / "bad.ssbscn")
        with open(q, "wb") as f:
            f.write(data)
        with pytest.raises(EngineError, match=why):
            scenario.Image.open(q)


def test_compile_rejects_what_the_device_record_cannot_hold(tmp_path):
    rd, trip = str(tmp_path / "rd.csv"), str(tmp_path / "trip.csv")
    with open(rd, "w") as f:
        f.write("R,0,0,1,40,500,1\nR,1,1,0,40,500,1\n")
    with open(trip, "w") as f:
        f.write("TP,1,0,4294967296,0,1\n")                                     # departure >= 2^31
    with pytest.raises(EngineError, match="trip file line 1"):
        scenario.compile_traffic_csv(rd, trip, str(tmp_path / "x.ssbscn"))
    with open(rd, "w") as f:
        f.write("R,0,0,1,40,500,1\n\n")                                        # an LP line without a road
    with open(trip, "w") as f:
        f.write("TP,1,0,5,0,1\n")
    with pytest.raises(EngineError, match="road file line 2"):
        scenario.compile_traffic_csv(rd, trip, str(tmp_path / "x.ssbscn"))
    assert not os.path.exists(str(tmp_path / "x.ssbscn")) and not os.path.exists(str(tmp_path / "x.ssbscn.tmp"))


def test_written_csv_is_what_the_reference_binary_reads(tmp_path):
    """The unmodified reference (oracle/_ref/TrafficSim, where built) run on files written by ssb_traffic_write_csv
    prints the ring's golden lines: the writer's format is the reference's."""
    exe = os.path.join(ROOT, "oracle", "_ref", "TrafficSim")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/TrafficSim not built (needs the reference tree)")
    sc, g = ring_scenario()
    sc.partition = np.zeros(sc.n_lp, dtype=np.int64)
    rd, trip, part = (str(tmp_path / n) for n in ("rd.csv", "trip.csv", "part"))
    scenario.write_traffic_csv(sc, rd, trip, part)
    lines = None
    for _ in range(4):                  # the reference has an intermittent start-up crash
        p = subprocess.run([exe, part, rd, trip], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120,
                           env=dict(os.environ, SCALESIM_NUM_THR="2"))
        if p.returncode == 0:
            lines = sorted(l for l in p.stdout.splitlines() if l.startswith("RE,"))
            break
    assert lines == g["lines"]


def test_native_readers_equal_the_line_readers_on_arbitrary_text(tmp_path):
    """Property test: on any text over the characters the formats use (and some they do not), the native readers and the
    per-line restatement of the reference's readers yield the same arrays — same line splitting ('\\n' only, as
    std::getline), same skipping of short lines, same atol reading."""
    import re
    from hypothesis import HealthCheck, assume, given, settings, strategies as st

    text = st.text(alphabet="0123456789,,,,;;--+ \tRTPx.\r\n\n", max_size=400)
    paths = [str(tmp_path / n) for n in ("a.txt",)]

    @settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(text, st.integers(-9, -1))      # negative: exactly that many pieces, however small the file
    def check(t, threads):
        assume(not re.search(r"\d{17,}", t))                 # beyond int64 the reference's atol is undefined
        with open(paths[0], "w", newline="", encoding="latin-1") as f:
            f.write(t)
        for kind in ("road", "trip", "part"):
            if kind == "road":
                nat = scenario.read_traffic_csv(road_csv=paths[0], threads=threads)
                py = dict(zip(("state_ids", "road_start", "road_dst", "road_speed", "road_len"), traffic.read_roads(paths[0])))
            elif kind == "trip":
                nat = scenario.read_traffic_csv(trip_csv=paths[0], threads=threads)
                py = dict(zip(("trip_id", "trip_dep", "route_start", "route_nodes"), traffic.read_trips(paths[0])))
            else:
                nat = scenario.read_traffic_csv(partition_file=paths[0], threads=threads)
                py = dict(part=traffic.read_partition(paths[0]))
            for k, v in py.items():
                assert np.array_equal(nat[k], v), (kind, k, t)

    check()
