"""Host side of the exact-differential history store (reference ``--diff_init`` / ``--diff_repeat``).

The reference keeps three LevelDBs per rank, ``<store_dir>/<sim_id>/{event,cancel,state}<rank>``
(leveldb_store.hpp:75-102), filled just before every fossil collection (logical_process.hpp:187-203,
queue.hpp:179-201, 304-323) and read back lazily per touched LP (queue.hpp:213-241, 325-331). Here the history of a
run is ONE pair of flat arrays drained from the device's committed log (``Engine.drain_history``):

    rec[i]   ssb_raw_record   the committed event (key = receive_time << 32 | id, dst, src, send_time, payload, tag)
    sent[i]  ssb_sent_record  the event it produced (the sent-log entry); dst 0xffffffff = none

which is the logical content of all three reference stores:

    event  DB : one record per committed event,       key (dst, receive_time, id)            queue.hpp:179-189
    cancel DB : one record per event that produced    key (dst, receive_time, id)            queue.hpp:191-201
                something (the produced event's send_time is the input's receive_time, same id — the app convention)
    state  DB : the same keys as cancel + one (lp, -1, -1) initial version per LP that        queue.hpp:304-323,
                executed an event (only LPs the scheduler dequeued are stored)                 runner.hpp:573-583

``key60`` renders the reference's 60-byte LevelDB key (leveldb_store.hpp:336-368) so that stores can be compared
record by record. The value bytes of the reference (Boost binary archives) are not reproduced (unpinned).

``HistoryWriter`` streams the store to its file from a background thread while the simulation runs
(``Engine.run_stream_history`` hands it the chunks from the engine's pinned staging buffers).
"""
from __future__ import annotations

import os
import threading

import numpy as np

from .capi import RAW_RECORD_DTYPE, SENT_RECORD_DTYPE

MAGIC = b"SSBHIST3"
NONE = 0xFFFFFFFF


def key60(lp: int, time: int, ident: int) -> bytes:
    """leveldb_store::to_key (leveldb_store.hpp:336-368): three 19-character right-aligned zero-padded decimal
    fields, each followed by NUL. Negative numbers render as the reference renders them ('...0-1')."""
    return b"".join(str(int(x)).rjust(19, "0").encode() + b"\0" for x in (lp, time, ident))


class HistoryStore:
    def __init__(self, rec: np.ndarray, sent: np.ndarray, state_lps: np.ndarray | None = None,
                 pool: np.ndarray | None = None):
        self.rec = np.ascontiguousarray(rec, dtype=RAW_RECORD_DTYPE)
        self.sent = np.ascontiguousarray(sent, dtype=SENT_RECORD_DTYPE)
        if self.rec.size != self.sent.size:
            raise ValueError("history arrays differ in length")
        #: LPs that had an initial state on this rank (each contributes the (-1,-1) state record)
        self.state_lps = np.ascontiguousarray(state_lps if state_lps is not None else np.zeros(0), dtype=np.int64)
        #: model data the recorded events refer to (TrafficSim: the route pool); travels with the history because a
        #: --diff_repeat run is not given the trip file
        self.pool = np.ascontiguousarray(pool if pool is not None else np.zeros(0), dtype=np.int64)

    # ---- logical records of the three reference stores, as sorted (lp, time, id) rows ----
    def _keys(self, mask=None) -> np.ndarray:
        r = self.rec if mask is None else self.rec[mask]
        k = np.stack([r["dst"].astype(np.int64), (r["key"] >> np.uint64(32)).astype(np.int64),
                      (r["key"] & np.uint64(0xFFFFFFFF)).astype(np.int64)], axis=1)
        return k[np.lexsort((k[:, 2], k[:, 1], k[:, 0]))]

    def event_keys(self) -> np.ndarray:
        return self._keys()

    def cancel_keys(self) -> np.ndarray:
        return self._keys(self.sent["dst"] != NONE)

    def active_lps(self) -> np.ndarray:
        """LPs of this rank that executed a committed event. The reference stores / fossil-collects only LPs its
        scheduler ever dequeued (runner::clear walks scheduler::active_lp, runner.hpp:573-583,
        process_scheduler.hpp:55-67), so an LP no event ever reached has no record in any of its stores — not even its
        initial state. (An LP reached only by events at or beyond finish_time is dequeued or not depending on the
        reference's thread timing: its (lp, -1, -1) record is not deterministic there and is not produced here.)"""
        return np.unique(self.rec["dst"].astype(np.int64))

    def state_keys(self) -> np.ndarray:
        k = self.cancel_keys()
        lps = np.intersect1d(self.state_lps, self.active_lps())
        init = np.stack([lps, np.full(lps.size, -1), np.full(lps.size, -1)], axis=1)
        k = np.concatenate([init.astype(np.int64), k]) if init.size else k
        return k[np.lexsort((k[:, 2], k[:, 1], k[:, 0]))]

    # ---- files ----
    @staticmethod
    def path(store_dir: str, sim_id, rank: int) -> str:
        return os.path.join(store_dir, str(sim_id), f"history{rank}.ssb")

    def save(self, store_dir: str, sim_id, rank: int = 0) -> str:
        w = HistoryWriter(self.path(store_dir, sim_id, rank), self.state_lps, self.pool)
        w.append(self.rec, self.sent)
        return w.close()

    def save_async(self, store_dir: str, sim_id, rank: int = 0) -> threading.Thread:
        """Write the store from a background thread (the arrays must stay untouched until ``join()``)."""
        t = threading.Thread(target=self.save, args=(store_dir, sim_id, rank), daemon=True)
        t.start()
        return t

    @classmethod
    def load(cls, store_dir: str, sim_id, rank: int = 0) -> "HistoryStore":
        p = cls.path(store_dir, sim_id, rank)
        size = os.path.getsize(p)
        with open(p, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{p}: not a scalesim_b200 history store")
            ns, npool = (int(x) for x in np.fromfile(f, dtype="<i8", count=2))
            if ns < 0 or npool < 0 or (ns + npool) * 8 > size:
                raise ValueError(f"{p}: header counts do not fit the file")
            lps = np.fromfile(f, dtype="<i8", count=ns)
            pool = np.fromfile(f, dtype="<i8", count=npool)
            recs, sents = [], []
            while True:
                hdr = np.fromfile(f, dtype="<i8", count=1)
                if hdr.size != 1:
                    raise ValueError(f"{p}: truncated (no closing chunk)")
                n = int(hdr[0])
                if n == 0:
                    break
                if n < 0 or n * 48 > size:
                    raise ValueError(f"{p}: bad chunk size")
                r = np.fromfile(f, dtype=RAW_RECORD_DTYPE, count=n)
                t = np.fromfile(f, dtype=SENT_RECORD_DTYPE, count=n)
                if r.size != n or t.size != n:
                    raise ValueError(f"{p}: truncated chunk")
                recs.append(r)
                sents.append(t)
            total = np.fromfile(f, dtype="<i8", count=1)
        rec = np.concatenate(recs) if recs else np.zeros(0, dtype=RAW_RECORD_DTYPE)
        sent = np.concatenate(sents) if sents else np.zeros(0, dtype=SENT_RECORD_DTYPE)
        if lps.size != ns or pool.size != npool or total.size != 1 or int(total[0]) != rec.size:
            raise ValueError(f"{p}: truncated")
        return cls(rec, sent, lps, pool)


class HistoryWriter:
    """Appends (records, sent-log entries) chunks to a history file from a background thread while the simulation
    runs (``Engine.run_stream_history``): host memory stays bounded by the queue depth, where the reference keeps
    every put in one in-memory WriteBatch until finish() (leveldb_store.hpp:132-141).
    File layout (same as include/scalesim/simulation/runner.hpp): "SSBHIST3", n_state, n_pool, state LPs, model data
    pool, then chunks {n, rec[n], sent[n]}, closed by n = 0 and the total record count."""

    def __init__(self, path: str, state_lps=None, pool=None, queue_depth: int = 4):
        import queue
        self.path = path
        os.makedirs(os.path.dirname(path), exist_ok=True)
        self.f = open(path + ".tmp", "wb")
        lps = np.ascontiguousarray(state_lps if state_lps is not None else np.zeros(0), dtype="<i8")
        pl = np.ascontiguousarray(pool if pool is not None else np.zeros(0), dtype="<i8")
        self.f.write(MAGIC)
        np.array([lps.size, pl.size], dtype="<i8").tofile(self.f)
        lps.tofile(self.f)
        pl.tofile(self.f)
        self.total = 0
        self.error = None
        self.q = queue.Queue(maxsize=queue_depth)
        self.thr = threading.Thread(target=self._loop, daemon=True)
        self.thr.start()

    def _loop(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            try:
                rec, sent = item
                np.array([rec.size], dtype="<i8").tofile(self.f)
                rec.tofile(self.f)
                sent.tofile(self.f)
                self.total += rec.size
            except Exception as exc:      # reported by close()
                self.error = exc

    def append(self, rec: np.ndarray, sent: np.ndarray, copy: bool = True):
        """Queue one chunk (copied: a staging-buffer view is reused by the engine as soon as the sink returns)."""
        if rec.size == 0:
            return
        r = np.array(rec, dtype=RAW_RECORD_DTYPE, copy=copy)
        t = np.array(sent, dtype=SENT_RECORD_DTYPE, copy=copy)
        self.q.put((r, t))

    def close(self) -> str:
        self.q.put(None)
        self.thr.join()
        if self.error is not None:
            self.f.close()
            raise self.error
        np.array([0, self.total], dtype="<i8").tofile(self.f)
        self.f.close()
        os.replace(self.path + ".tmp", self.path)
        return self.path
