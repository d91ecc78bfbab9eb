"""ctypes binding of oracle/libtw_oracle.so — TEST INFRASTRUCTURE (the checker). Never imported by the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORD_DTYPE = np.dtype([("id", "<i8"), ("src", "<i8"), ("send_time", "<i8"), ("dst", "<i8"), ("recv_time", "<i8"),
                         ("tag", "<i8")])

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "libtw_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
        L = C.CDLL(path)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        L.two_digest_records.argtypes = [vp, i64, vp]
        L.two_lp_new.restype = vp
        L.two_lp_new.argtypes = [i64]
        L.two_lp_free.argtypes = [vp]
        L.two_lp_buffer.argtypes = [vp, i64, i64, i64, i64, i64, i32]
        L.two_lp_flush.restype = i64
        L.two_lp_flush.argtypes = [vp, vp, i64]
        L.two_lp_dequeue.argtypes = [vp, vp]
        L.two_lp_direct_insert.argtypes = [vp, i64, i64, i64, i64, i64]
        L.two_lp_local_time.argtypes = [vp, vp]
        L.two_lp_set_cancel.argtypes = [vp, i64, i64, i64, i64, i64]
        L.two_lp_init_state.argtypes = [vp, i64]
        L.two_lp_update_state.argtypes = [vp, i64, i64, i64]
        L.two_lp_get_state.restype = i64
        L.two_lp_get_state.argtypes = [vp]
        L.two_lp_num_states.restype = i64
        L.two_lp_num_states.argtypes = [vp]
        L.two_lp_std_out.restype = i64
        L.two_lp_std_out.argtypes = [vp, i64, i64, vp, i64]
        L.two_lp_clear.argtypes = [vp, i64, i64]
        L.two_lp_size_ev.restype = i64
        L.two_lp_size_ev.argtypes = [vp]
        L.two_lp_size_can.restype = i64
        L.two_lp_size_can.argtypes = [vp]
        L.two_phold_tables.argtypes = [i64, dbl, dbl, i64, vp, vp]
        L.two_sim_phold.restype = vp
        L.two_sim_phold.argtypes = [i64, i64, dbl, dbl, i64]
        L.two_sim_traffic_files.restype = vp
        L.two_sim_traffic_files.argtypes = [C.c_char_p, C.c_char_p]
        L.two_sim_traffic_arrays.restype = vp
        L.two_sim_traffic_arrays.argtypes = [i64, vp, vp, vp, vp, i64, vp, vp, vp, vp]
        L.two_sim_free.argtypes = [vp]
        L.two_sim_set_finish_time.argtypes = [vp, i64]
        L.two_sim_set_stateful.argtypes = [vp, i32]
        L.two_sim_run_timewarp.restype = i64
        L.two_sim_run_timewarp.argtypes = [vp, i32, i32, i32]
        L.two_sim_run_closure.restype = i64
        L.two_sim_run_closure.argtypes = [vp, i32, i32]
        L.two_sim_run_sequential.restype = i64
        L.two_sim_run_sequential.argtypes = [vp, i32]
        L.two_sim_num_committed.restype = i64
        L.two_sim_num_committed.argtypes = [vp]
        L.two_sim_committed.restype = i64
        L.two_sim_committed.argtypes = [vp, vp, i64]
        L.two_sim_digest.argtypes = [vp, vp]
        L.two_sim_stats.argtypes = [vp, vp]
        L.two_sim_final_states.restype = i64
        L.two_sim_final_states.argtypes = [vp, vp, vp, i64]
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class Sim:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: could not create simulation (missing input file?)")
        self.h = C.c_void_p(handle)

    @classmethod
    def phold(cls, num_lp, num_init, ratio, lam, seed=1, stateful=False):
        s = cls(lib().two_sim_phold(num_lp, num_init, ratio, lam, seed))
        if stateful:
            lib().two_sim_set_stateful(s.h, 1)
        return s

    @classmethod
    def traffic_files(cls, road_csv, trip_csv):
        return cls(lib().two_sim_traffic_files(road_csv.encode(), trip_csv.encode()))

    @classmethod
    def traffic_scenario(cls, sc):
        """sc: scalesim_b200.traffic.Scenario whose state_ids are 0..n_lp-1 in order (synthetic grid)."""
        a = [np.ascontiguousarray(x, dtype=np.int64) for x in
             (sc.road_start, sc.road_dst, sc.road_speed, sc.road_len, sc.trip_id, sc.trip_dep, sc.route_start,
              sc.route_nodes)]
        s = cls(lib().two_sim_traffic_arrays(sc.state_ids.size, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]),
                                             sc.trip_id.size, _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7])))
        return s

    def set_finish_time(self, t):
        lib().two_sim_set_finish_time(self.h, t)

    def run_timewarp(self, num_sched=4, gsync_interval=100, switch_lp_interval=1):
        return lib().two_sim_run_timewarp(self.h, num_sched, gsync_interval, switch_lp_interval)

    def run_closure(self, threads=4, keep_records=True):
        return lib().two_sim_run_closure(self.h, threads, 1 if keep_records else 0)

    def run_sequential(self, keep_records=True):
        return lib().two_sim_run_sequential(self.h, 1 if keep_records else 0)

    def committed(self) -> np.ndarray:
        n = lib().two_sim_num_committed(self.h)
        out = np.empty(n, dtype=RECORD_DTYPE)
        k = lib().two_sim_committed(self.h, _p(out), n)
        return out[:k]

    def digest(self):
        d = np.zeros(4, dtype=np.uint64)
        lib().two_sim_digest(self.h, _p(d))
        return tuple(int(x) for x in d)

    def stats(self):
        s = np.zeros(4, dtype=np.int64)
        lib().two_sim_stats(self.h, _p(s))
        return dict(processed=int(s[0]), cancels=int(s[1]), gvt_rounds=int(s[2]), remote=int(s[3]))

    def final_states(self, n_lp):
        ids = np.empty(n_lp, dtype=np.int64)
        words = np.empty((n_lp, 2), dtype=np.uint32)
        k = lib().two_sim_final_states(self.h, _p(ids), _p(words), n_lp)
        return ids[:k], words[:k]

    def close(self):
        if self.h:
            lib().two_sim_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def digest_records(rec: np.ndarray):
    d = np.zeros(4, dtype=np.uint64)
    r = np.ascontiguousarray(rec)
    lib().two_digest_records(_p(r), r.size, _p(d))
    return tuple(int(x) for x in d)


def phold_closure_digest(num_lp, num_init, ratio, lam, threads=4):
    s = Sim.phold(num_lp, num_init, ratio, lam)
    s.run_closure(threads, keep_records=False)
    d = s.digest()
    s.close()
    return d


def records_to_lines(rec):
    suffix = ("", ",START", ",END")
    return [f"RE,{r['id']},{r['src']},{r['send_time']},{r['dst']},{r['recv_time']}{suffix[int(r['tag'])]}" for r in rec]


class LP:
    """LP-level oracle object (restated lp<App>)."""

    def __init__(self, lp_id=0):
        self.h = C.c_void_p(lib().two_lp_new(lp_id))

    def buffer(self, id, recv, send=0, dst=0, src=0, cancel=False):
        lib().two_lp_buffer(self.h, id, src, dst, recv, send, 1 if cancel else 0)

    def flush(self):
        out = np.empty(4096, dtype=RECORD_DTYPE)
        n = lib().two_lp_flush(self.h, _p(out), out.size)
        return out[:n]

    def dequeue(self):
        out = np.empty(1, dtype=RECORD_DTYPE)
        ok = lib().two_lp_dequeue(self.h, _p(out))
        return out[0] if ok else None

    def direct_insert(self, id, recv, send=0, dst=0, src=0):
        lib().two_lp_direct_insert(self.h, id, src, dst, recv, send)

    def local_time(self):
        t = np.zeros(2, dtype=np.int64)
        lib().two_lp_local_time(self.h, _p(t))
        return int(t[0]), int(t[1])

    def set_cancel(self, id, recv, send, dst=0, src=0):
        lib().two_lp_set_cancel(self.h, id, src, dst, recv, send)

    def init_state(self, v):
        lib().two_lp_init_state(self.h, v)

    def update_state(self, v, t, i):
        lib().two_lp_update_state(self.h, v, t, i)

    def get_state(self):
        return lib().two_lp_get_state(self.h)

    def num_states(self):
        return lib().two_lp_num_states(self.h)

    def std_out(self, t, i):
        out = np.empty(4096, dtype=RECORD_DTYPE)
        n = lib().two_lp_std_out(self.h, t, i, _p(out), out.size)
        return out[:n]

    def clear(self, t, i):
        lib().two_lp_clear(self.h, t, i)

    def size_ev(self):
        return lib().two_lp_size_ev(self.h)

    def size_can(self):
        return lib().two_lp_size_can(self.h)

    def __del__(self):
        try:
            lib().two_lp_free(self.h)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# exact-differential repeat (R13): which recorded events a set of injected events forces to be re-executed.
# Plain-Python restatement of the reference's cascade (logical_process.hpp:129-157, queue.hpp:82-108, 213-225):
# an LP touched at key m reloads / rolls back everything it holds with key >= m, cancels and re-sends the outputs of
# those events, and every such output touches its destination LP at the output's key. The fixpoint over "earliest
# touch per LP" is the deterministic part of the reference's --diff_repeat output; on top of it the reference prints
# history that an LP reloaded because of a transient optimistic message (timing dependent — see make_golden_diff.py).
# ------------------------------------------------------------------------------------------------
def parse_line(l):
    f = l.split(",")
    return (int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]))      # id, src, send, dst, recv


def repeat_closure(init_lines, new_lines, deletes=()):
    """Lines of the repeat run = new_lines + every init line at an LP whose key (recv, id) is >= the LP's earliest touch.
    deletes = (lp, time, id) of deleted events (DE / RE queries, runner.hpp:246-276): the LP is touched at that key,
    the event itself disappears, and — eventq::delete_can, queue.hpp:238-241 — its output is neither cancelled nor
    redone, so the delete does not propagate through it. A deleted event comes BACK when the cascade of another query
    makes its sender re-execute the event that produced it: the sender's anti-message finds nothing, the re-sent
    positive is inserted as a new event (seen in the reference on the 12x12 grid)."""
    import heapq
    deleted = {(int(lp), (int(t), int(i))) for lp, t, i in deletes}
    by_lp = {}
    succ = {}
    pred = {}
    for l in init_lines:
        i, src, send, dst, recv = parse_line(l)
        by_lp.setdefault(dst, []).append(((recv, i), l))
        if (src, send) != (dst, recv):                   # PHOLD's initial events name themselves as their source
            succ[(i, src, send)] = (dst, (recv, i))      # the output of the event (id, at LP src, key (send, id))
            pred[(dst, (recv, i))] = (src, (send, i))
    for v in by_lp.values():
        v.sort()
    touch = {}
    heap = []
    for l in new_lines:
        i, src, send, dst, recv = parse_line(l)
        heapq.heappush(heap, ((recv, i), dst))
    for lp, key in deleted:
        heapq.heappush(heap, (key, lp))
    revived = set()
    while True:
        while heap:
            key, lp = heapq.heappop(heap)
            old = touch.get(lp)
            if old is not None and old <= key:
                continue
            touch[lp] = key
            for k, l in by_lp.get(lp, []):
                if k >= key and (old is None or k < old) and ((lp, k) not in deleted or (lp, k) in revived):
                    s = succ.get((k[1], lp, k[0]))
                    if s is not None:
                        heapq.heappush(heap, (s[1], s[0]))
        again = False
        for d in deleted - revived:
            p = pred.get(d)
            if p is not None and p[0] in touch and touch[p[0]] <= p[1] and p not in (deleted - revived):
                revived.add(d)                            # its sender re-executed the producing event: re-sent
                s = succ.get((d[1][1], d[0], d[1][0]))
                if s is not None:
                    heapq.heappush(heap, (s[1], s[0]))
                again = True
        if not again:
            break
    gone = deleted - revived
    out = list(new_lines)
    for lp, m in touch.items():
        out += [l for k, l in by_lp.get(lp, []) if k >= m and (lp, k) not in gone]
    return sorted(out)
