"""Developer tool: a multi-GPU run on the host emulation — the partitions are engines of ONE process, each driven by
its own host thread (ctypes releases the GIL during ssb_run), their mailboxes exchanged by address (cuda_emu.h's
in-process "IPC"). The kernels of the peer-memory exchange (k_route to a peer's mailbox, k_gvt, the receive pass,
sequence stamps, anti-messages across partitions) run unchanged except for k_gvt's per-partition threads, which the
emulation unrolls. Logic check only; multi-GPU parity is claimed from runs on real B200s.

    run_partitions(make_engine, world) -> (combined digest, [Stats per rank], engines closed)
"""
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import emu_plugin  # noqa: F401,E402
from scalesim_b200 import dist  # noqa: E402


def connect(engines, p2p=True):
    """Peer-memory exchange: ssb_comm_export on every engine, then ssb_comm_import of all handles on every engine.
    p2p=False: the NCCL exchange mode (ssb_comm_unique_id / ssb_comm_init) over the emulation's in-process NCCL."""
    if not p2p:
        from scalesim_b200 import Engine
        uid = Engine.comm_unique_id()
        for e in engines:
            e.comm_init(uid)
        return
    handles = b"".join(e.comm_export() for e in engines)
    for e in engines:
        e.comm_import(handles)


def run_all(engines, call=lambda e: e.run()):
    """call(engine) on one thread per engine, concurrently; returns the results in rank order; re-raises the first error."""
    out = [None] * len(engines)
    err = [None] * len(engines)

    def work(i):
        try:
            out[i] = call(engines[i])
        except BaseException as e:      # noqa: BLE001
            err[i] = e

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(engines))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    # a partition that stops with its own error (e.g. SSB_ERR_CAPACITY) leaves its peers waiting at the GVT exchange
    # until they give up with SSB_ERR_COMM: report the cause, not the consequence
    errs = [e for e in err if e is not None]
    if errs:
        raise next((e for e in errs if getattr(e, "status", None) != 6), errs[0])
    return out


def run_partitions(make_engine, world, keep_engines=False, graph=True, sync_interval=1, p2p=True):
    """make_engine(rank, world) -> loaded Engine. graph: the interpreted graph loop (ssb_engine::run_graph_emulated) or
    the stream path; sync_interval: local windows per global synchronisation (graph loop only, ssb_set_sync_interval).
    Returns (digest of the union, stats per rank[, engines])."""
    if not graph:
        os.environ["SSB_NO_GRAPH"] = "1"
    try:
        engines = [make_engine(r, world) for r in range(world)]
    finally:
        os.environ.pop("SSB_NO_GRAPH", None)
    connect(engines, p2p)
    if sync_interval > 1:
        for e in engines:
            e.set_sync_interval(sync_interval)
    stats = run_all(engines)
    digest = dist.combine_digests([e.digest() for e in engines])
    if keep_engines:
        return digest, stats, engines
    for e in engines:
        e.close()
    return digest, stats
