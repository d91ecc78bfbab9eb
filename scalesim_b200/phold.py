"""Host-side mirror of ScaleSim's PHOLD application (reference src/phold/phold.hpp).

Everything here is input preparation on the host; the handler itself runs on the GPU
(``PholdModel`` in csrc/ssb_models.cuh).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import (EVENT_DTYPE, MODEL_PHOLD, MODEL_PHOLD_STATEFUL, Config, Engine, load_library)

EFFECTIVE_DECIMAL = 100000            # phold.hpp:54
LOOK_AHEAD = int(0.1 * EFFECTIVE_DECIMAL)   # phold.hpp:55
RAND_TABLE_SIZE = 10_000_000          # phold.hpp:56
FINISH_TIME = 5 * EFFECTIVE_DECIMAL   # phold.hpp:138
RANDOM_SEED = 1                       # main_phold.cc:26


def build_tables(remote_ratio: float, lam: float, seed: int = RANDOM_SEED, n: int = RAND_TABLE_SIZE):
    """LATENCY_TABLE / REMOTE_COM_TABLE (phold.hpp:144-164): libstdc++ streams, built by the C++ helper."""
    lib = load_library()
    lat = np.empty(n, dtype=np.int64)
    rem = np.empty(n, dtype=np.int32)
    rc = lib.ssb_phold_build_tables(seed, lam, remote_ratio, n, C.c_void_p(lat.ctypes.data), C.c_void_p(rem.ctypes.data))
    if rc != 0:
        raise RuntimeError("ssb_phold_build_tables failed")
    return lat, rem


def partition(num_lp: int, world: int) -> np.ndarray:
    """phold.hpp:176-189: round robin, lp % rank_size."""
    return np.arange(num_lp, dtype=np.int64) % world


def init_events(num_lp: int, num_init_msg: int, latency: np.ndarray, rank: int = 0, world: int = 1,
                out: np.ndarray | None = None) -> np.ndarray:
    """Initial events owned by ``rank`` (phold.hpp:197-209 followed by the shuffle of runner.hpp:131-147)."""
    lib = load_library()
    if out is None:
        out = np.empty(num_init_msg, dtype=EVENT_DTYPE)
    n = lib.ssb_phold_init_events(num_lp, num_init_msg, C.c_void_p(latency.ctypes.data), latency.size, 0, 1,
                                  C.c_void_p(out.ctypes.data), out.size)
    if n < 0:
        raise RuntimeError("ssb_phold_init_events failed")
    ev = out[:n]
    if world > 1:
        ev = ev[(ev["dst"] % world) == rank]
    return ev


def init_states(num_lp: int, rank: int = 0, world: int = 1, stateful: bool = False):
    """phold.hpp:216-225: State{id} for every LP of this rank. Device layout: 2 words (id as u64);
    the stateful test model starts from {count = 0, chain = lp id}."""
    ids = np.arange(rank, num_lp, world, dtype=np.int64)
    words = np.zeros((ids.size, 2), dtype=np.uint32)
    if stateful:
        words[:, 1] = ids.astype(np.uint32)
    else:
        words[:, 0] = (ids & 0xFFFFFFFF).astype(np.uint32)
        words[:, 1] = (ids >> 32).astype(np.uint32)
    return ids, words


def default_capacity(num_init_msg: int, world: int = 1):
    per = (num_init_msg + world - 1) // world
    return int(per * 1.6) + (1 << 16)


def make_engine(num_lp: int, num_init_msg: int, remote_ratio: float, lam: float, *, tables=None,
                bucket_ticks: int = LOOK_AHEAD, window_buckets: int = 1, keep_records: bool = False,
                rank: int = 0, world: int = 1, device: int = 0, stateful: bool = False,
                max_events: int | None = None, max_batch: int | None = None, max_committed: int | None = None,
                finish_time: int = FINISH_TIME, rounds_per_sync: int = 4, load: bool = True, flags: int = 0,
                max_snapshots: int = 1 << 16) -> Engine:
    """Build a PHOLD engine the way runner<phold>::init does (runner.hpp:79-176)."""
    lat, rem = tables if tables is not None else build_tables(remote_ratio, lam)
    per = (num_init_msg + world - 1) // world
    if max_events is None:
        max_events = default_capacity(num_init_msg, world)
    if max_batch is None:
        max_batch = max(1 << 14, min(per, int(per * 0.25 * window_buckets)) + (1 << 14))
    if max_committed is None:
        max_committed = int(per * 6) if keep_records else 0
    cfg = Config(model=MODEL_PHOLD_STATEFUL if stateful else MODEL_PHOLD, n_lp=num_lp, finish_time=finish_time,
                 bucket_ticks=bucket_ticks, window_buckets=window_buckets, max_events=max_events,
                 max_batch=max_batch, max_committed=max_committed, rank=rank, world=world, device=device,
                 rounds_per_sync=rounds_per_sync, flags=flags, max_snapshots=max_snapshots)
    eng = Engine(cfg)
    eng.set_partition(None)
    eng.set_model_data(0, lat)
    eng.set_model_data(1, rem)
    eng.set_model_data(2, np.array([LOOK_AHEAD, num_lp], dtype=np.int64))
    ids, words = init_states(num_lp, rank, world, stateful)
    eng.load_states(ids, words)
    if load:
        eng.load_events(init_events(num_lp, num_init_msg, lat, rank, world))
    return eng
