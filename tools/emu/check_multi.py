"""Developer tool / CPU logic check: a few multi-partition runs on the host emulation (tools/emu/multi.py), printed as
one JSON object. Run by tests/test_emu_logic.py in a subprocess (the emulation is selected per process).

    tools/emu/build.sh && python tools/emu/check_multi.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import multi  # noqa: E402  (selects the emulation)
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
from scalesim_b200 import phold, traffic  # noqa: E402


def main():
    out = {}
    tables = phold.build_tables(0.5, 1.0)
    want = O.phold_closure_digest(1000, 16000, 0.5, 1.0)
    for world, window in ((2, 1), (2, 3), (3, 2), (4, 1)):
        d, st = multi.run_partitions(lambda r, w: phold.make_engine(1000, 16000, 0.5, 1.0, tables=tables, rank=r, world=w,
                                                                    window_buckets=window), world)
        out[f"phold_w{world}_win{window}"] = dict(equal=d == want, committed=sum(s.committed for s in st),
                                                  remote=sum(s.remote_sent for s in st), rollbacks=sum(s.rollbacks for s in st),
                                                  antis=sum(s.antis_sent for s in st))
    # state-mutating model: records, digest and the final LP states against the sequential DES
    s = O.Sim.phold(300, 3000, 0.5, 1.0, stateful=True)
    s.run_sequential()
    ids, words = s.final_states(300)
    d, st, engines = multi.run_partitions(lambda r, w: phold.make_engine(300, 3000, 0.5, 1.0, tables=tables, rank=r, world=w,
                                                                         window_buckets=4, stateful=True), 3, keep_engines=True)
    got = np.zeros_like(words)
    for r, e in enumerate(engines):
        mine = ids % 3 == r
        got[mine] = e.read_states(ids[mine])
        e.close()
    out["stateful_w3"] = dict(equal=d == s.digest(), states_equal=bool(np.array_equal(got, words)),
                              rollbacks=sum(x.rollbacks for x in st))
    # TrafficSim: 32x32 torus in row stripes, and a random network with a random partition table
    sc = traffic.synthetic_grid(32, 32, 8000, seed=1)
    sc.partition = traffic.stripe_partition(32, 32, 4)
    o = O.Sim.traffic_scenario(sc)
    o.run_closure(2)
    for world, window in ((2, 1), (4, 3)):
        d, st = multi.run_partitions(lambda r, w: traffic.make_engine(sc, rank=r, world=w, bucket_ticks=11, window_buckets=window), world)
        out[f"grid_w{world}_win{window}"] = dict(equal=d == o.digest(), remote=sum(x.remote_sent for x in st),
                                                 rollbacks=sum(x.rollbacks for x in st))
    # several local windows per global synchronisation (ssb_set_sync_interval; graph loop): optimism across partitions
    for world, k in ((2, 2), (4, 8)):
        d, st = multi.run_partitions(lambda r, w: traffic.make_engine(sc, rank=r, world=w, bucket_ticks=11), world, sync_interval=k)
        out[f"grid_w{world}_sync{k}"] = dict(equal=d == o.digest(), rollbacks=sum(x.rollbacks for x in st),
                                             rounds=st[0].rounds, exchanges=st[0].exchanges)
    d, st = multi.run_partitions(lambda r, w: phold.make_engine(1000, 16000, 0.5, 1.0, tables=tables, rank=r, world=w), 3, sync_interval=4)
    out["phold_w3_sync4"] = dict(equal=d == want, rollbacks=sum(s.rollbacks for s in st), rounds=st[0].rounds, exchanges=st[0].exchanges)
    # the stream path (kernels launched round by round) instead of the graph loop
    d, st = multi.run_partitions(lambda r, w: phold.make_engine(1000, 16000, 0.5, 1.0, tables=tables, rank=r, world=w, window_buckets=2), 2, graph=False)
    out["phold_w2_stream"] = dict(equal=d == want, rollbacks=sum(s.rollbacks for s in st))
    # the NCCL exchange mode (ncclSend / ncclRecv + all-reduce min; here the emulation's in-process NCCL)
    d, st = multi.run_partitions(lambda r, w: phold.make_engine(1000, 16000, 0.5, 1.0, tables=tables, rank=r, world=w, window_buckets=3), 3, p2p=False)
    out["phold_w3_nccl"] = dict(equal=d == want, rollbacks=sum(s.rollbacks for s in st), remote=sum(s.remote_sent for s in st))
    d, st = multi.run_partitions(lambda r, w: traffic.make_engine(sc, rank=r, world=w, bucket_ticks=11, window_buckets=2), 4, p2p=False)
    out["grid_w4_nccl"] = dict(equal=d == o.digest(), rollbacks=sum(x.rollbacks for x in st), remote=sum(x.remote_sent for x in st))
    # every partition started from ONE scenario image (ssb_scenario_load keeps the rank's share)
    import scalesim_b200 as ssb
    from scalesim_b200 import scenario
    im = scenario.traffic_image(sc)

    def from_image(r, w):
        e = ssb.Engine(ssb.Config(model=ssb.MODEL_TRAFFIC, n_lp=sc.n_lp, finish_time=traffic.FINISH_TIME, bucket_ticks=11, window_buckets=2,
                                  max_events=sc.trip_id.size * 2 + (1 << 16), max_batch=sc.trip_id.size + (1 << 14), rank=r, world=w))
        e.load_scenario(im)
        return e
    d, st = multi.run_partitions(from_image, 4)
    out["grid_w4_from_image"] = dict(equal=d == o.digest(), remote=sum(x.remote_sent for x in st), rollbacks=sum(x.rollbacks for x in st))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
