"""Multi-GPU plumbing: one process per GPU. torch.distributed is used only for the rendezvous (mailbox IPC handles or
the NCCL unique id) and for combining per-rank results; the data path (cross-partition events, the GVT min-reduction)
is inside libscalesim_b200.so: peer-memory mailboxes over NVLink by default, NCCL send/recv + all-reduce on request."""
from __future__ import annotations

import numpy as np

from .capi import Engine


def init_engine_comm(engine: Engine, rank: int, world: int, p2p: bool = True):
    """Rank 0 creates the NCCL unique id, everyone receives it and joins the communicator (ssb_comm_init). With
    ``p2p`` the engines also swap the CUDA IPC handles of their mailboxes, so cross-partition events are written
    straight into the destination GPU's memory over NVLink (ssb_comm_export / ssb_comm_import); otherwise they
    travel by ncclSend/ncclRecv."""
    if world <= 1:
        return
    import torch.distributed as dist
    if p2p:
        # peer-memory exchange: events and GVT candidates cross through the mailboxes; no NCCL communicator needed
        handles = [None] * world
        dist.all_gather_object(handles, engine.comm_export())
        engine.comm_import(b"".join(handles))
        return
    obj = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    engine.comm_init(obj[0])


def combine_digests(digests):
    """Digest of the union of the partitions' committed multisets: count / sums add (mod 2^64), the xor word xors."""
    m = (1 << 64) - 1
    d0 = d1 = d2 = d3 = 0
    for d in digests:
        d0 = (d0 + d[0]) & m
        d1 = (d1 + d[1]) & m
        d2 ^= d[2]
        d3 = (d3 + d[3]) & m
    return (d0, d1, d2, d3)


def all_gather_digest(digest, world: int):
    if world <= 1:
        return tuple(digest)
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, tuple(int(x) for x in digest))
    return combine_digests(out)


def max_over_ranks(x: float, world: int) -> float:
    if world <= 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: int, world: int) -> int:
    if world <= 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def owner_of(lp_ids: np.ndarray, world: int, partition: np.ndarray | None = None) -> np.ndarray:
    """Partition that owns each LP: the application's table (values taken modulo world, as the reference's
    road_read does, traffic_sim.hpp:377) or PHOLD's round robin lp % world (phold.hpp:181-186)."""
    lp_ids = np.asarray(lp_ids, dtype=np.int64)
    if partition is None:
        return lp_ids % world
    return np.asarray(partition, dtype=np.int64)[lp_ids] % world
