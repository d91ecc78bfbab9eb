"""ctypes binding of include/scalesim_b200.h (the C ABI of libscalesim_b200.so)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

MODEL_PHOLD = 1
MODEL_TRAFFIC = 2
MODEL_PHOLD_STATEFUL = 3
ABI_VERSION = 2

#: host mirror of ``ssb_event`` (8 x int64)
EVENT_DTYPE = np.dtype([
    ("id", "<i8"), ("src", "<i8"), ("dst", "<i8"), ("recv_time", "<i8"), ("send_time", "<i8"),
    ("p0", "<i8"), ("p1", "<i8"), ("flags", "<i8"),
])
#: host mirror of ``ssb_record`` (6 x int64): one committed "RE,..." line
RECORD_DTYPE = np.dtype([
    ("id", "<i8"), ("src", "<i8"), ("send_time", "<i8"), ("dst", "<i8"), ("recv_time", "<i8"), ("tag", "<i8"),
])


#: host mirror of ``ssb_raw_record`` (32 bytes): the committed log as it sits in HBM
RAW_RECORD_DTYPE = np.dtype([
    ("key", "<u8"), ("dst", "<u4"), ("src", "<u4"), ("send_time", "<u4"), ("p0", "<u4"), ("p1", "<u4"), ("tag", "<u4"),
])
#: host mirror of ``ssb_packed_record`` (24 bytes): the six fields of one committed line, 32 bits each
PACKED_RECORD_DTYPE = np.dtype([("id", "<u4"), ("src", "<u4"), ("send_time", "<u4"), ("dst", "<u4"), ("recv_time", "<u4"),
                                ("tag", "<u4")])
#: host mirror of ``ssb_sent_record`` (16 bytes): the event a committed event produced (dst 0xffffffff = none)
SENT_RECORD_DTYPE = np.dtype([("dst", "<u4"), ("id", "<u4"), ("recv_time", "<u4"), ("reserved", "<u4")])
FLAG_DEBUG_CHECKS = 1
FLAG_KERNEL_TIMING = 2
FLAG_HISTORY = 4


def events_to_raw(ev: np.ndarray) -> np.ndarray:
    """ssb_event (64 bytes) -> the compact 32-byte layout that ssb_load_events_raw takes."""
    raw = np.zeros(ev.shape, dtype=RAW_RECORD_DTYPE)
    raw["key"] = (ev["recv_time"].astype(np.uint64) << np.uint64(32)) | ev["id"].astype(np.uint64)
    raw["dst"] = ev["dst"]
    raw["src"] = np.where(ev["src"] < 0, 0xFFFFFFFF, ev["src"]).astype(np.uint32)
    raw["send_time"] = ev["send_time"]
    raw["p0"] = ev["p0"]
    raw["p1"] = ev["p1"]
    raw["tag"] = ev["flags"] & 1
    return raw


def packed_to_records(p: np.ndarray) -> np.ndarray:
    """ssb_packed_record -> ssb_record (widening; src 0xffffffff -> -1)."""
    rec = np.empty(p.shape, dtype=RECORD_DTYPE)
    for f in ("id", "send_time", "dst", "recv_time", "tag"):
        rec[f] = p[f]
    src = p["src"].astype(np.int64)
    src[p["src"] == 0xFFFFFFFF] = -1
    rec["src"] = src
    return rec


def raw_to_records(raw: np.ndarray) -> np.ndarray:
    """ssb_raw_record -> ssb_record (widening; src 0xffffffff -> -1)."""
    rec = np.empty(raw.shape, dtype=RECORD_DTYPE)
    rec["id"] = (raw["key"] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    rec["recv_time"] = (raw["key"] >> np.uint64(32)).astype(np.int64)
    src = raw["src"].astype(np.int64)
    src[raw["src"] == 0xFFFFFFFF] = -1
    rec["src"] = src
    rec["send_time"] = raw["send_time"]
    rec["dst"] = raw["dst"]
    rec["tag"] = raw["tag"]
    return rec


class EngineError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"scalesim_b200 status {status}: {message}")
        self.status = status
        self.message = message


class _Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("window_buckets", C.c_int32),
        ("n_lp", C.c_int64), ("finish_time", C.c_int64), ("bucket_ticks", C.c_int64),
        ("max_events", C.c_int64), ("max_batch", C.c_int64), ("max_committed", C.c_int64),
        ("max_snapshots", C.c_int64), ("rounds_per_sync", C.c_int32), ("flags", C.c_int32),
    ]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "rounds", "processed", "rollbacks", "undone", "antis_sent", "annihilated", "duplicates",
        "committed", "beyond_finish", "remote_sent", "gvt_time", "gvt_id", "pool_high_water", "log_dropped",
    )] + [("loop_ms", C.c_double), ("fix_rounds", C.c_int64), ("fix_turns", C.c_int64),
                                                                                ("exchanges", C.c_int64)]


@dataclass
class Config:
    model: int
    n_lp: int
    finish_time: int
    bucket_ticks: int
    window_buckets: int = 1
    max_events: int = 1 << 20
    max_batch: int = 1 << 18
    max_committed: int = 0
    max_snapshots: int = 1 << 16
    rounds_per_sync: int = 4
    device: int = 0
    rank: int = 0
    world: int = 1
    flags: int = 0


@dataclass
class Stats:
    rounds: int = 0
    processed: int = 0
    rollbacks: int = 0
    undone: int = 0
    antis_sent: int = 0
    annihilated: int = 0
    duplicates: int = 0
    committed: int = 0
    beyond_finish: int = 0
    remote_sent: int = 0
    gvt_time: int = 0
    gvt_id: int = 0
    pool_high_water: int = 0
    log_dropped: int = 0
    loop_ms: float = 0.0
    fix_rounds: int = 0
    fix_turns: int = 0
    exchanges: int = 0


#: ssb_packed_sink / ssb_history_sink
PACKED_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)
HISTORY_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libscalesim_b200.so")


_lib = None


def load_library() -> C.CDLL:
    """Load libscalesim_b200.so (built in-tree by ``__graft_entry__.build()``). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(-1, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.ssb_abi_version.restype = i32
    lib.ssb_last_error.restype = C.c_char_p
    lib.ssb_last_error.argtypes = [vp]
    lib.ssb_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.ssb_destroy.argtypes = [vp]
    lib.ssb_destroy.restype = None
    lib.ssb_set_partition.argtypes = [vp, vp, i64]
    lib.ssb_set_model_data.argtypes = [vp, i32, vp, C.c_size_t]
    lib.ssb_state_words.argtypes = [vp]
    lib.ssb_load_states.argtypes = [vp, vp, vp, i64]
    lib.ssb_read_states.argtypes = [vp, vp, vp, i64]
    lib.ssb_load_events.argtypes = [vp, vp, i64]
    lib.ssb_run.argtypes = [vp, C.POINTER(_Stats)]
    lib.ssb_step.argtypes = [vp, C.POINTER(i32)]
    lib.ssb_get_stats.argtypes = [vp, C.POINTER(_Stats)]
    lib.ssb_committed_digest.argtypes = [vp, vp]
    lib.ssb_drain_committed.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.ssb_drain_committed_raw.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.ssb_load_events_raw.argtypes = [vp, vp, i64]
    lib.ssb_drain_committed_packed.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.ssb_run_drain_packed.argtypes = [vp, vp, i64, C.POINTER(i64), C.POINTER(_Stats)]
    lib.ssb_drain_history.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    lib.ssb_run_stream_packed.argtypes = [vp, PACKED_SINK, vp, C.POINTER(_Stats)]
    lib.ssb_run_stream_history.argtypes = [vp, HISTORY_SINK, vp, C.POINTER(_Stats)]
    lib.ssb_load_history.argtypes = [vp, vp, vp, i64]
    lib.ssb_set_state_from.argtypes = [vp, i64, i64, vp]
    lib.ssb_reset.argtypes = [vp]
    lib.ssb_set_flags.argtypes = [vp, C.c_int32]
    lib.ssb_kernel_times.argtypes = [vp, vp, vp, vp, i32]
    lib.ssb_comm_unique_id.argtypes = [vp]
    lib.ssb_comm_init.argtypes = [vp, vp]
    lib.ssb_comm_export.argtypes = [vp, vp]
    lib.ssb_comm_import.argtypes = [vp, vp, C.c_int32]
    lib.ssb_set_sync_interval.argtypes = [vp, C.c_int32]
    lib.ssb_phold_build_tables.argtypes = [i64, C.c_double, C.c_double, i64, vp, vp]
    lib.ssb_phold_init_events.argtypes = [i64, i64, vp, i64, i64, i64, vp, i64]
    lib.ssb_phold_init_events.restype = i64
    lib.ssb_traffic_read.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i32, vp]
    lib.ssb_traffic_free.argtypes = [vp]
    lib.ssb_traffic_free.restype = None
    lib.ssb_traffic_write_csv.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p]
    lib.ssb_scenario_save.argtypes = [C.c_char_p, vp]
    lib.ssb_scenario_open.argtypes = [C.c_char_p, vp]
    lib.ssb_scenario_close.argtypes = [vp]
    lib.ssb_scenario_close.restype = None
    lib.ssb_scenario_load.argtypes = [vp, vp]
    lib.ssb_traffic_compile.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i64, i32, C.c_char_p]
    lib.ssb_get_config.argtypes = [vp, C.POINTER(_Config)]
    if lib.ssb_abi_version() != ABI_VERSION:
        raise EngineError(-1, "libscalesim_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Engine:
    """One Time Warp engine = one partition on one GPU (``ssb_engine``)."""

    def __init__(self, cfg: Config):
        self._lib = load_library()
        self._h = C.c_void_p()
        c = _Config(ABI_VERSION, cfg.model, cfg.device, cfg.rank, cfg.world, cfg.window_buckets, cfg.n_lp,
                    cfg.finish_time, cfg.bucket_ticks, cfg.max_events, cfg.max_batch, cfg.max_committed,
                    cfg.max_snapshots, cfg.rounds_per_sync, cfg.flags)
        rc = self._lib.ssb_create(C.byref(c), C.byref(self._h))
        if rc != 0:
            msg = self._lib.ssb_last_error(None).decode()
            self._h = C.c_void_p()
            raise EngineError(rc, msg)
        self.cfg = cfg

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.ssb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise EngineError(rc, self._lib.ssb_last_error(self._h).decode())

    # -- setup --
    def set_partition(self, lp_to_part=None):
        if lp_to_part is None:
            self._check(self._lib.ssb_set_partition(self._h, None, 0))
        else:
            a = np.ascontiguousarray(lp_to_part, dtype=np.int64)
            self._check(self._lib.ssb_set_partition(self._h, _ptr(a), a.size))

    def set_model_data(self, slot: int, data: np.ndarray):
        a = np.ascontiguousarray(data)
        self._check(self._lib.ssb_set_model_data(self._h, slot, _ptr(a), a.nbytes))

    @property
    def state_words(self) -> int:
        return self._lib.ssb_state_words(self._h)

    def load_states(self, lp_ids, words):
        ids = np.ascontiguousarray(lp_ids, dtype=np.int64)
        w = np.ascontiguousarray(words, dtype=np.uint32).reshape(ids.size, self.state_words)
        self._check(self._lib.ssb_load_states(self._h, _ptr(ids), _ptr(w), ids.size))

    def read_states(self, lp_ids) -> np.ndarray:
        ids = np.ascontiguousarray(lp_ids, dtype=np.int64)
        w = np.empty((ids.size, self.state_words), dtype=np.uint32)
        self._check(self._lib.ssb_read_states(self._h, _ptr(ids), _ptr(w), ids.size))
        return w

    def load_events(self, events: np.ndarray):
        assert events.dtype == EVENT_DTYPE
        a = np.ascontiguousarray(events)
        self._check(self._lib.ssb_load_events(self._h, _ptr(a), a.size))

    def load_scenario(self, image):
        """Partition, model data and this rank's states and events from a scenario image
        (``scalesim_b200.scenario.Image``): ssb_scenario_load."""
        self._check(self._lib.ssb_scenario_load(self._h, C.byref(image.c)))

    def config(self) -> Config:
        """The configuration the engine was created with, read back through ssb_get_config."""
        c = _Config()
        self._check(self._lib.ssb_get_config(self._h, C.byref(c)))
        return Config(**{n: getattr(c, n) for n, _ in _Config._fields_ if n != "abi_version"})

    def load_events_ptr(self, ptr: int, n: int):
        """Load ``n`` ssb_event records from a raw host pointer (e.g. a pinned torch tensor)."""
        self._check(self._lib.ssb_load_events(self._h, C.c_void_p(ptr), n))

    # -- run --
    def run(self) -> Stats:
        s = _Stats()
        self._check(self._lib.ssb_run(self._h, C.byref(s)))
        return self._stats(s)

    def step(self) -> bool:
        done = C.c_int(0)
        self._check(self._lib.ssb_step(self._h, C.byref(done)))
        return bool(done.value)

    def stats(self) -> Stats:
        s = _Stats()
        self._check(self._lib.ssb_get_stats(self._h, C.byref(s)))
        return self._stats(s)

    @staticmethod
    def _stats(s: _Stats) -> Stats:
        return Stats(**{n: getattr(s, n) for n, _ in _Stats._fields_})

    def digest(self):
        d = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.ssb_committed_digest(self._h, _ptr(d)))
        return tuple(int(x) for x in d)

    def drain_committed(self, cap: int | None = None) -> np.ndarray:
        if cap is None:
            cap = max(1, self.stats().committed)
        out = np.empty(cap, dtype=RECORD_DTYPE)
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_committed(self._h, _ptr(out), cap, C.byref(n)))
        return out[: n.value]

    def drain_committed_ptr(self, ptr: int, cap: int) -> int:
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_committed(self._h, C.c_void_p(ptr), cap, C.byref(n)))
        return n.value

    def drain_committed_raw(self, cap: int | None = None) -> np.ndarray:
        """Committed records in the compact 32-byte device layout (``ssb_raw_record``)."""
        if cap is None:
            cap = max(1, self.stats().committed)
        out = np.empty(cap, dtype=RAW_RECORD_DTYPE)
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_committed_raw(self._h, _ptr(out), cap, C.byref(n)))
        return out[: n.value]

    def drain_committed_raw_ptr(self, ptr: int, cap: int) -> int:
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_committed_raw(self._h, C.c_void_p(ptr), cap, C.byref(n)))
        return n.value

    def load_events_raw(self, raw: np.ndarray):
        raw = np.ascontiguousarray(raw, dtype=RAW_RECORD_DTYPE)
        self._check(self._lib.ssb_load_events_raw(self._h, _ptr(raw), raw.size))

    def load_events_raw_ptr(self, ptr: int, n: int):
        self._check(self._lib.ssb_load_events_raw(self._h, C.c_void_p(ptr), n))

    def drain_committed_packed(self, cap: int | None = None) -> np.ndarray:
        """Committed records in the 24-byte packed layout (``ssb_packed_record``)."""
        if cap is None:
            cap = max(1, self.stats().committed)
        out = np.empty(cap, dtype=PACKED_RECORD_DTYPE)
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_committed_packed(self._h, _ptr(out), cap, C.byref(n)))
        return out[: n.value]

    def drain_committed_packed_ptr(self, ptr: int, cap: int) -> int:
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_committed_packed(self._h, C.c_void_p(ptr), cap, C.byref(n)))
        return n.value

    def run_drain_packed_ptr(self, ptr: int, cap: int):
        """Run the simulation while the committed records stream to host memory (ssb_run_drain_packed).
        Returns (stats, number of records)."""
        s = _Stats()
        n = C.c_int64(0)
        self._check(self._lib.ssb_run_drain_packed(self._h, C.c_void_p(ptr), cap, C.byref(n), C.byref(s)))
        return self._stats(s), n.value

    def run_drain_packed(self, cap: int):
        out = np.empty(cap, dtype=PACKED_RECORD_DTYPE)
        st, n = self.run_drain_packed_ptr(out.ctypes.data, cap)
        return st, out[:n]

    def run_stream_packed(self, on_chunk) -> Stats:
        """ssb_run_stream_packed: ``on_chunk(records)`` gets every committed record (``PACKED_RECORD_DTYPE`` view of the
        engine's pinned staging buffer — copy what you keep) while the window loop runs."""
        def sink(_user, rec, n):
            buf = (C.c_char * (n * PACKED_RECORD_DTYPE.itemsize)).from_address(rec)
            return int(on_chunk(np.frombuffer(buf, dtype=PACKED_RECORD_DTYPE, count=n)) or 0)
        s = _Stats()
        self._check(self._lib.ssb_run_stream_packed(self._h, PACKED_SINK(sink), None, C.byref(s)))
        return self._stats(s)

    def run_stream_history(self, on_chunk) -> Stats:
        """ssb_run_stream_history (--diff_init): ``on_chunk(rec, sent)`` gets every committed event with its sent-log
        entry, chunk by chunk, while the window loop runs (views of the pinned staging buffers)."""
        def sink(_user, rec, sent, n):
            a = (C.c_char * (n * RAW_RECORD_DTYPE.itemsize)).from_address(rec)
            b = (C.c_char * (n * SENT_RECORD_DTYPE.itemsize)).from_address(sent)
            return int(on_chunk(np.frombuffer(a, dtype=RAW_RECORD_DTYPE, count=n),
                                np.frombuffer(b, dtype=SENT_RECORD_DTYPE, count=n)) or 0)
        s = _Stats()
        self._check(self._lib.ssb_run_stream_history(self._h, HISTORY_SINK(sink), None, C.byref(s)))
        return self._stats(s)

    def drain_history(self, cap: int | None = None):
        """--diff_init: every committed event (``ssb_raw_record``) with its sent-log entry (``ssb_sent_record``).
        Needs ``Config.flags & FLAG_HISTORY`` and ``max_committed > 0``."""
        if cap is None:
            cap = max(1, self.stats().committed)
        rec = np.empty(cap, dtype=RAW_RECORD_DTYPE)
        sent = np.empty(cap, dtype=SENT_RECORD_DTYPE)
        n = C.c_int64(0)
        self._check(self._lib.ssb_drain_history(self._h, _ptr(rec), _ptr(sent), cap, C.byref(n)))
        return rec[: n.value], sent[: n.value]

    def load_history(self, rec: np.ndarray, sent: np.ndarray):
        """--diff_repeat: put a recorded history back as processed events (nothing is executed)."""
        rec = np.ascontiguousarray(rec, dtype=RAW_RECORD_DTYPE)
        sent = np.ascontiguousarray(sent, dtype=SENT_RECORD_DTYPE)
        if rec.size != sent.size:
            raise ValueError("history arrays differ in length")
        self._check(self._lib.ssb_load_history(self._h, _ptr(rec), _ptr(sent), rec.size))

    def set_state_from(self, lp_id: int, time: int, words):
        """--diff_repeat SC query: LP ``lp_id`` gets the state ``words``; its recorded history from ``time`` on is invalid."""
        w = np.ascontiguousarray(words, dtype=np.uint32).reshape(self.state_words)
        self._check(self._lib.ssb_set_state_from(self._h, lp_id, time, _ptr(w)))

    def reset(self):
        self._check(self._lib.ssb_reset(self._h))

    def set_flags(self, flags: int):
        self._check(self._lib.ssb_set_flags(self._h, flags))

    def kernel_times(self) -> dict:
        """{kernel name: (total device ms, launches)}; needs FLAG_KERNEL_TIMING."""
        cap = 16
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        ln = (C.c_int64 * cap)()
        k = self._lib.ssb_kernel_times(self._h, names, ms, ln, cap)
        return {names[i].decode(): (ms[i], ln[i]) for i in range(k)}

    # -- multi-GPU --
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.ssb_comm_unique_id(buf)
        if rc != 0:
            raise EngineError(rc, lib.ssb_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == 128
        self._check(self._lib.ssb_comm_init(self._h, C.c_char_p(unique_id)))

    def comm_export(self) -> bytes:
        """64-byte CUDA IPC handle of this engine's mailbox (peer-memory exchange)."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.ssb_comm_export(self._h, buf))
        return buf.raw

    def set_sync_interval(self, windows: int):
        """Local windows per global synchronisation (peer-memory exchange, ``run()`` only); see ssb_set_sync_interval."""
        self._check(self._lib.ssb_set_sync_interval(self._h, windows))

    def comm_import(self, handles: bytes):
        world = len(handles) // 64
        self._check(self._lib.ssb_comm_import(self._h, C.c_char_p(handles), world))


# ---------------------------------------------------------------------------------------
# committed records -> the reference's output
# ---------------------------------------------------------------------------------------
def records_to_lines(rec: np.ndarray) -> list[str]:
    """``RE,id,src,send,dst,recv[,START|,END]`` (reference sim_obj.hpp:66-77), one per record."""
    suffix = ("", ",START", ",END")
    return [f"RE,{r['id']},{r['src']},{r['send_time']},{r['dst']},{r['recv_time']}{suffix[int(r['tag'])]}" for r in rec]


_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z ^ (z >> np.uint64(30))
    z = z * _M1
    z = z ^ (z >> np.uint64(27))
    z = z * _M2
    return z ^ (z >> np.uint64(31))


def digest_records(rec: np.ndarray):
    """Order-independent digest of a multiset of committed records; same function as the device
    digest (ssb_device.cuh record_hash) and the oracle's (tw_oracle.cc record_hash)."""
    with np.errstate(over="ignore"):
        h = np.full(rec.shape, 0x9E3779B97F4A7C15, dtype=np.uint64)
        for f in ("id", "src", "send_time", "dst", "recv_time", "tag"):
            h = _mix64(h ^ rec[f].astype(np.int64).view(np.uint64))
        s1 = int(np.add.reduce(h, dtype=np.uint64)) if h.size else 0
        x = int(np.bitwise_xor.reduce(h)) if h.size else 0
        s2 = int(np.add.reduce(_mix64(h + np.uint64(0xD1B54A32D192ED03)), dtype=np.uint64)) if h.size else 0
    return (int(rec.size), s1, x, s2)
