"""world_size-2 tests of the multi-partition HOST logic on CPU (gloo): sharding of the initial events and states by
owner, the rendezvous helpers, and the combination of per-partition results. The device exchange itself (NCCL) is
covered by tests/test_gpu_multi.py on a 2-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scalesim_b200 as ssb
from scalesim_b200 import phold, traffic
from fixtures import golden, golden_lines, ring_scenario


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _lines_to_records(lines):
    rec = np.zeros(len(lines), dtype=ssb.RECORD_DTYPE)
    for i, l in enumerate(lines):
        f = l.split(",")
        rec[i] = (int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), {"START": 1, "END": 2}.get(f[6], 0) if len(f) > 6 else 0)
    return rec


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. every rank takes the committed records whose destination LP it owns (what its GPU would commit)
        rec = _lines_to_records(golden_lines("phold_100_1600"))
        mine = rec[ssb.dist.owner_of(rec["dst"], world) == rank]
        total = ssb.dist.sum_over_ranks(int(mine.size), world)
        combined = ssb.dist.all_gather_digest(ssb.digest_records(mine), world)
        # 2. initial events: shards are disjoint, cover everything, and every event goes to the owner of its LP
        lat, _ = phold.build_tables(0.1, 1.0, n=4096)
        ev = phold.init_events(100, 1600, lat, rank, world)
        ids = [None] * world
        dist.all_gather_object(ids, ev["id"].tolist())
        owner_ok = bool(((ev["dst"] % world) == rank).all())
        # 3. states
        sid, words = phold.init_states(100, rank, world)
        # 4. max over ranks
        mx = ssb.dist.max_over_ranks(float(rank + 1), world)
        # 5. object broadcast used for the NCCL unique id
        obj = [b"x" * 128 if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        q.put((rank, total, combined, sorted(sum(ids, [])), owner_ok, sid.tolist(), mx, obj[0]))
    finally:
        dist.destroy_process_group()


def test_two_partitions_host_logic():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = golden("phold_100_1600")
    for rank, total, combined, all_ids, owner_ok, sid, mx, blob in res:
        assert total == g["count"]
        assert combined == g["digest"]                    # union of the partitions == the reference's output
        assert all_ids == list(range(1600))               # disjoint cover (phold_test.cc:45-104)
        assert owner_ok
        assert sid == list(range(rank, 100, world))
        assert mx == 2.0
        assert blob == b"x" * 128


def test_combine_digests_matches_whole():
    rec = _lines_to_records(golden_lines("phold_100_1600"))
    parts = [rec[i::3] for i in range(3)]
    assert ssb.dist.combine_digests([ssb.digest_records(p) for p in parts]) == ssb.digest_records(rec)


def test_traffic_owner_uses_partition_table():
    sc, _ = ring_scenario()
    assert ssb.dist.owner_of(np.arange(10), 4, sc.partition).tolist() == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]
    assert ssb.dist.owner_of(np.arange(10), 2, sc.partition).tolist() == [0, 1, 0, 1, 0, 1, 0, 1, 0, 1]
