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
    state  DB : the same keys as cancel + one (lp, -1, -1) initial version per LP             queue.hpp:304-323

``key60`` renders the reference's 60-byte LevelDB key (leveldb_store.hpp:336-368) so that stores can be compared
record by record. The value bytes of the reference (Boost binary archives) are not reproduced (unpinned).

Files are written by a background thread from the (pinned) host arrays while the caller goes on (``save_async``).
"""
from __future__ import annotations

import os
import threading

import numpy as np

from .capi import RAW_RECORD_DTYPE, SENT_RECORD_DTYPE

MAGIC = b"SSBHIST2"
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

    def state_keys(self) -> np.ndarray:
        k = self.cancel_keys()
        init = np.stack([self.state_lps, np.full(self.state_lps.size, -1), np.full(self.state_lps.size, -1)], axis=1)
        k = np.concatenate([init.astype(np.int64), k]) if init.size else k
        return k[np.lexsort((k[:, 2], k[:, 1], k[:, 0]))]

    # ---- files ----
    @staticmethod
    def path(store_dir: str, sim_id, rank: int) -> str:
        return os.path.join(store_dir, str(sim_id), f"history{rank}.ssb")

    def save(self, store_dir: str, sim_id, rank: int = 0) -> str:
        p = self.path(store_dir, sim_id, rank)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p + ".tmp", "wb") as f:
            f.write(MAGIC)
            np.array([self.rec.size, self.state_lps.size, self.pool.size], dtype="<i8").tofile(f)
            self.rec.tofile(f)
            self.sent.tofile(f)
            self.state_lps.astype("<i8").tofile(f)
            self.pool.astype("<i8").tofile(f)
        os.replace(p + ".tmp", p)
        return p

    def save_async(self, store_dir: str, sim_id, rank: int = 0) -> threading.Thread:
        """Write the store from a background thread (the arrays must stay untouched until ``join()``)."""
        t = threading.Thread(target=self.save, args=(store_dir, sim_id, rank), daemon=True)
        t.start()
        return t

    @classmethod
    def load(cls, store_dir: str, sim_id, rank: int = 0) -> "HistoryStore":
        p = cls.path(store_dir, sim_id, rank)
        with open(p, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{p}: not a scalesim_b200 history store")
            n, ns, npool = (int(x) for x in np.fromfile(f, dtype="<i8", count=3))
            rec = np.fromfile(f, dtype=RAW_RECORD_DTYPE, count=n)
            sent = np.fromfile(f, dtype=SENT_RECORD_DTYPE, count=n)
            lps = np.fromfile(f, dtype="<i8", count=ns)
            pool = np.fromfile(f, dtype="<i8", count=npool)
        if rec.size != n or sent.size != n or lps.size != ns or pool.size != npool:
            raise ValueError(f"{p}: truncated")
        return cls(rec, sent, lps, pool)
