"""Input at scale: the native TrafficSim readers and the scenario image (``ssb_traffic_*`` / ``ssb_scenario_*`` of
include/scalesim_b200.h; implementation csrc/scenario_host.cc).

The reference parses its scenario as text on every rank, one ``getline`` + ``boost::split`` + ``atol`` per line and one
``shared_ptr`` object per event (src/trafficsim/traffic_sim.hpp:345-451); a 10 M-trip scenario is ~3 GB of CSV.
Here the CSV is parsed once by all host cores over a mapped file (``read_traffic_csv``), and what an engine needs to
start is kept as ONE page-aligned binary file (``Image``) that later runs map and copy to the device without parsing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .capi import MODEL_PHOLD, MODEL_TRAFFIC, RAW_RECORD_DTYPE, EngineError, events_to_raw, load_library

SLOTS = 4


class _Tables(C.Structure):
    _fields_ = [
        ("n_part", C.c_int64), ("part", C.c_void_p),
        ("n_states", C.c_int64), ("state_ids", C.c_void_p), ("road_start", C.c_void_p),
        ("n_roads", C.c_int64), ("road_dst", C.c_void_p), ("road_speed", C.c_void_p), ("road_len", C.c_void_p),
        ("n_trips", C.c_int64), ("trip_id", C.c_void_p), ("trip_dep", C.c_void_p), ("route_start", C.c_void_p),
        ("n_nodes", C.c_int64), ("route_nodes", C.c_void_p),
    ]


class _Scenario(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("state_words", C.c_int32), ("n_lp", C.c_int64), ("finish_time", C.c_int64),
        ("n_part", C.c_int64), ("lp_to_part", C.c_void_p),
        ("n_states", C.c_int64), ("state_lps", C.c_void_p), ("states", C.c_void_p),
        ("slot_bytes", C.c_int64 * SLOTS), ("slot_data", C.c_void_p * SLOTS),
        ("n_events", C.c_int64), ("events", C.c_void_p),
        ("mapping", C.c_void_p), ("mapping_bytes", C.c_int64),
    ]


def _fail(lib, rc):
    raise EngineError(rc, lib.ssb_last_error(None).decode())


def _enc(path):
    return None if path is None else os.fsencode(path)


class _Owned:
    """The arrays ssb_traffic_read allocated: freed (ssb_traffic_free) when the last numpy view of them is gone."""

    def __init__(self, lib, tables):
        self.lib, self.tables = lib, tables

    def __del__(self):
        try:
            self.lib.ssb_traffic_free(C.byref(self.tables))
        except Exception:
            pass

    def view(self, ptr, n):
        if not n:
            return np.zeros(0, dtype=np.int64)
        buf = (C.c_char * (n * 8)).from_address(ptr)
        buf._owner = self                      # the view keeps the buffer, the buffer keeps the owner
        return np.frombuffer(buf, dtype=np.int64)


def read_traffic_csv(road_csv=None, trip_csv=None, partition_file=None, threads: int = 0) -> dict:
    """The three readers of the reference's TrafficSim (traffic_reader::graph_read / road_read / trip_read,
    traffic_sim.hpp:345-451) run natively on ``threads`` host threads (0 = all cores). Returns flat int64 arrays
    (views of the parser's own output, no copy): part, state_ids, road_start, road_dst, road_speed, road_len, trip_id,
    trip_dep, route_start, route_nodes."""
    lib = load_library()
    t = _Tables()
    rc = lib.ssb_traffic_read(_enc(partition_file), _enc(road_csv), _enc(trip_csv), threads, C.byref(t))
    if rc:
        _fail(lib, rc)
    o = _Owned(lib, t)
    out = {"part": o.view(t.part, t.n_part) if partition_file else None}
    if road_csv:
        out.update(state_ids=o.view(t.state_ids, t.n_states), road_start=o.view(t.road_start, t.n_states + 1),
                   road_dst=o.view(t.road_dst, t.n_roads), road_speed=o.view(t.road_speed, t.n_roads),
                   road_len=o.view(t.road_len, t.n_roads))
    if trip_csv:
        out.update(trip_id=o.view(t.trip_id, t.n_trips), trip_dep=o.view(t.trip_dep, t.n_trips),
                   route_start=o.view(t.route_start, t.n_trips + 1), route_nodes=o.view(t.route_nodes, t.n_nodes))
    return out


def read_scenario(road_csv, trip_csv, partition_file=None, threads: int = 0):
    """``traffic.read_scenario`` with the native readers: a ``traffic.Scenario`` of the three files."""
    from . import traffic
    t = read_traffic_csv(road_csv, trip_csv, partition_file, threads)
    n_lp = int(max(t["state_ids"].max(initial=-1), t["road_dst"].max(initial=-1), t["route_nodes"].max(initial=-1))) + 1
    if t["part"] is not None:
        n_lp = max(n_lp, t["part"].size)
    return traffic.Scenario(n_lp, t["state_ids"], t["road_start"], t["road_dst"], t["road_speed"], t["road_len"],
                            t["trip_id"], t["trip_dep"], t["route_start"], t["route_nodes"], t["part"])


def write_traffic_csv(sc, road_csv=None, trip_csv=None, partition_file=None):
    """A ``traffic.Scenario`` as files in the reference's format (so that a generated scenario can be fed to the
    reference's own binary): ssb_traffic_write_csv."""
    lib = load_library()
    keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (
        sc.partition if sc.partition is not None else np.zeros(0, np.int64), sc.state_ids, sc.road_start, sc.road_dst,
        sc.road_speed, sc.road_len, sc.trip_id, sc.trip_dep, sc.route_start, sc.route_nodes)]
    p = [a.ctypes.data for a in keep]
    t = _Tables(keep[0].size, p[0], keep[1].size, p[1], p[2], keep[3].size, p[3], p[4], p[5],
                keep[6].size, p[6], p[7], p[8], keep[9].size, p[9])
    rc = lib.ssb_traffic_write_csv(C.byref(t), _enc(partition_file), _enc(road_csv), _enc(trip_csv))
    if rc:
        _fail(lib, rc)


def compile_traffic_csv(road_csv, trip_csv, image_path, partition_file=None, finish_time: int = 10800, threads: int = 0):
    """CSV files -> scenario image, natively (parse + pack + write): ssb_traffic_compile."""
    lib = load_library()
    rc = lib.ssb_traffic_compile(_enc(partition_file), _enc(road_csv), _enc(trip_csv), finish_time, threads, _enc(image_path))
    if rc:
        _fail(lib, rc)


class Image:
    """A scenario image: a read-only mapping of the file (``Image.open``) or arrays held by the caller (``Image.of``).
    ``engine.load_scenario(image)`` starts an engine of any rank / world from it."""

    def __init__(self):
        self.c = _Scenario()
        self._keep = []
        self._lib = load_library()
        self._mapped = False

    @classmethod
    def open(cls, path) -> "Image":
        im = cls()
        rc = im._lib.ssb_scenario_open(_enc(path), C.byref(im.c))
        if rc:
            _fail(im._lib, rc)
        im._mapped = True
        return im

    @classmethod
    def of(cls, model: int, n_lp: int, finish_time: int, state_lps, states, slots, events, partition=None) -> "Image":
        """``states``: uint32 [n_states, words]; ``slots``: list of arrays (ssb_set_model_data slot i, None = unused);
        ``events``: RAW_RECORD_DTYPE or EVENT_DTYPE records of every partition."""
        im = cls()
        c = im.c
        ids = np.ascontiguousarray(state_lps, dtype=np.int64)
        w = np.ascontiguousarray(states, dtype=np.uint32).reshape(ids.size, -1)
        ev = np.ascontiguousarray(events if events.dtype == RAW_RECORD_DTYPE else events_to_raw(events))
        im._keep += [ids, w, ev]
        c.model, c.state_words, c.n_lp, c.finish_time = model, w.shape[1], n_lp, finish_time
        c.n_states, c.state_lps, c.states = ids.size, ids.ctypes.data, w.ctypes.data
        c.n_events, c.events = ev.size, ev.ctypes.data
        if partition is not None:
            part = np.ascontiguousarray(partition, dtype=np.int64)
            im._keep.append(part)
            c.n_part, c.lp_to_part = part.size, part.ctypes.data
        for i, a in enumerate(slots):
            if a is None:
                continue
            a = np.ascontiguousarray(a)
            im._keep.append(a)
            c.slot_bytes[i], c.slot_data[i] = a.nbytes, a.ctypes.data
        return im

    def save(self, path):
        rc = self._lib.ssb_scenario_save(_enc(path), C.byref(self.c))
        if rc:
            _fail(self._lib, rc)

    def close(self):
        if self._mapped:
            self._lib.ssb_scenario_close(C.byref(self.c))
            self._mapped = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # views (valid while the image is open)
    model = property(lambda s: s.c.model)
    n_lp = property(lambda s: s.c.n_lp)
    finish_time = property(lambda s: s.c.finish_time)
    n_events = property(lambda s: s.c.n_events)
    n_states = property(lambda s: s.c.n_states)

    def _view(self, ptr, n, dtype):
        if not n:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype)

    @property
    def events(self):
        return self._view(self.c.events, self.c.n_events, RAW_RECORD_DTYPE)

    @property
    def partition(self):
        return self._view(self.c.lp_to_part, self.c.n_part, np.int64) if self.c.n_part else None

    @property
    def state_lps(self):
        return self._view(self.c.state_lps, self.c.n_states, np.int64)

    @property
    def states(self):
        return self._view(self.c.states, self.c.n_states * self.c.state_words, np.uint32).reshape(self.c.n_states, -1)

    def slot(self, i: int, dtype=np.int64):
        n = self.c.slot_bytes[i]
        return self._view(self.c.slot_data[i], n // np.dtype(dtype).itemsize, dtype) if n else None


def traffic_image(sc, finish_time: int = 10800) -> Image:
    """Image of a ``traffic.Scenario`` (same packing as ``traffic.make_engine``)."""
    from . import traffic
    ids, words, road_table = traffic.pack_states(sc)
    pool = sc.route_nodes.astype(np.int64) if sc.route_nodes.size else np.zeros(1, np.int64)
    return Image.of(MODEL_TRAFFIC, sc.n_lp, finish_time, ids, words,
                    [pool, road_table.reshape(-1) if road_table.size else None], traffic.pack_events(sc), sc.partition)


def phold_image(n_lp: int, n_init: int, remote_ratio: float, lam: float, tables=None) -> Image:
    """Image of a PHOLD scenario (tables + initial events of phold.hpp:144-209)."""
    from . import phold
    lat, rem = tables if tables is not None else phold.build_tables(remote_ratio, lam)
    ev = phold.init_events(n_lp, n_init, lat)
    params = np.array([phold.LOOK_AHEAD, n_lp], dtype=np.int64)
    ids = np.arange(n_lp, dtype=np.int64)
    words = np.stack([(ids & 0xFFFFFFFF).astype(np.uint32), (ids >> 32).astype(np.uint32)], axis=1)
    return Image.of(MODEL_PHOLD, n_lp, phold.FINISH_TIME, ids, words, [lat, rem, params], ev)


def ingest_measure(width: int = 1024, height: int = 1024, n_trips: int = 500_000, workdir: str | None = None,
                   bucket_ticks: int = 11, device: int = 0, run: bool = True) -> dict:
    """Scenario ingestion end to end, timed (host wall clock): a generated torus scenario written in the reference's
    CSV format -> native readers -> scenario image -> engine (ssb_scenario_load), next to the array path
    (``traffic.make_engine``) on the same scenario; with ``run`` both engines simulate and their committed digests must
    be equal. Used by ``bench.py`` (key ``ingest``) and by tests/test_gpu_scenario.py."""
    import shutil
    import tempfile
    import time

    from . import traffic
    from .capi import Config, Engine

    own = workdir is None
    d = tempfile.mkdtemp(prefix="ssb_ingest_") if own else workdir
    out = {"workload": f"TrafficSim {width}x{height} torus grid, {n_trips} trips as CSV in the reference's format"}
    try:
        sc = traffic.synthetic_grid(width, height, n_trips, seed=1)
        rd, trip, img = (os.path.join(d, n) for n in ("rd.csv", "trip.csv", "scenario.ssbscn"))
        t = time.perf_counter()
        write_traffic_csv(sc, rd, trip)
        out["write_csv_s"] = round(time.perf_counter() - t, 4)
        out["csv_bytes"] = os.path.getsize(rd) + os.path.getsize(trip)
        t = time.perf_counter()
        tabs = read_traffic_csv(rd, trip)
        out["native_read_s"] = round(time.perf_counter() - t, 4)
        out["native_read_gb_per_s"] = round(out["csv_bytes"] / 1e9 / max(out["native_read_s"], 1e-9), 3)
        out["host_threads"] = os.cpu_count()
        same = all(np.array_equal(tabs[k], getattr(sc, k)) for k in (
            "state_ids", "road_start", "road_dst", "road_speed", "road_len", "trip_id", "trip_dep", "route_start", "route_nodes"))
        out["read_equals_generator"] = bool(same)
        del tabs
        t = time.perf_counter()
        compile_traffic_csv(rd, trip, img, finish_time=traffic.FINISH_TIME)
        out["compile_to_image_s"] = round(time.perf_counter() - t, 4)
        out["image_bytes"] = os.path.getsize(img)

        kw = dict(max_events=int(n_trips * 1.5) + (1 << 16) + 4096 * 8, max_batch=max(1 << 14, n_trips // 2 + (1 << 14)))
        t = time.perf_counter()
        im = Image.open(img)
        out["open_image_s"] = round(time.perf_counter() - t, 5)
        eng = Engine(Config(model=MODEL_TRAFFIC, n_lp=im.n_lp, finish_time=im.finish_time, bucket_ticks=bucket_ticks,
                            device=device, rounds_per_sync=8, **kw))
        t = time.perf_counter()
        eng.load_scenario(im)
        out["load_image_s"] = round(time.perf_counter() - t, 4)      # mapping -> HBM: partition, tables, states, events
        im.close()
        digest_image = None
        if run:
            st = eng.run()
            digest_image = eng.digest()
            out["committed"] = st.committed
        eng.close()
        t = time.perf_counter()
        eng2 = traffic.make_engine(sc, bucket_ticks=bucket_ticks, device=device, rounds_per_sync=8, **kw)
        out["array_path_pack_and_load_s"] = round(time.perf_counter() - t, 4)   # numpy packing + per-array ABI calls
        if run:
            eng2.run()
            out["digest_equal"] = bool(eng2.digest() == digest_image)
        eng2.close()
        return out
    finally:
        if own:
            shutil.rmtree(d, ignore_errors=True)
