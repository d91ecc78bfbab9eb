#!/usr/bin/env python
"""bench.py — committed events/s of the Time Warp hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete PHOLD simulation (BASELINE.json config 2 per GPU: 1M LPs, 16M initial events,
10 % remote, lambda 1.0, finish 500 000 ticks => 72 823 220 committed events at N=1) through the C ABI.
  value : committed events / s of the window loop with the initial events already in HBM (device time, CUDA events,
          max over ranks) — the reference's "Simulation elapsed time" excludes init the same way (runner.hpp:71-73)
  e2e   : the same metric through the public host API with HOST buffers: reset + H2D of the initial events from
          pinned memory (ssb_load_events_raw, 32 B per event) + window loop + digest + D2H of every committed record
          into pinned memory (ssb_run_drain_packed: 24 B per record, overlapped with the loop), wall clock
  --impl reference : the reference's own CPU implementation (oracle/_ref, unmodified sources; else the oracle port)
          on a bounded PHOLD sample, same metric/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "committed events/sec, PHOLD 1M LPs"
UNIT = "events/s"
B_ALG = 380.0   # algorithmic bytes per committed PHOLD event (SURVEY.md §8d / DESIGN.md §5)
# per-kernel split of the 380 B: execute (pop read 48 + table gathers 12 + state read 8 + snapshot write 24 +
# sent-log write 24 + emit write 48), route (bucket read 48 + pending write 48), commit + fossil (read 48 + 40-byte
# result record + fossil key reads 32)
B_ALG_KERNEL = {"k_exec": 164.0, "k_route": 96.0, "k_commit": 120.0}
FULL_COMMITTED_1M = 72_823_220
# kernels per round inside the window-loop graph: plan, commit, exec, route = 4; N > 1 adds k_gvt and the two
# receive-side routes; a round with an irregular LP turn (ssb_stats.fix_rounds) adds the IF node's six kernels:
# k_fix_count, k_fix_segs, k_fix_scatter, k_fix_sort, k_fix_turn, k_route(anti).
def launches(st, world):
    return int(st.rounds * (4 + (3 if world > 1 else 0)) + st.fix_rounds * 6)


# ------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.lines = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference itself (oracle/_ref) or the oracle port, on a bounded sample
# ------------------------------------------------------------------------------------------------
# bounded PHOLD samples of the config-2 shape (16 events per LP, 0.1 remote): (LPs, initial events, committed events,
# events/s relative to the 10 000-LP sample — the reference's commit / fossil passes walk every active LP per GVT round,
# so its rate drops as the LP count grows)
SAMPLES = [(100000, 1600000, 7275345, 0.60), (30000, 480000, 2184583, 0.85), (10000, 160000, 728183, 1.0)]


def cpu_threads():
    n = os.cpu_count() or 1
    return max(1, min(8, n - 1))


def run_reference_once(n_lp, n_init, threads, timeout=300):
    """Unmodified reference PHOLD (single MPI rank, `threads` worker threads, -O0 as its CMakeLists demands).
    Returns (committed, loop_seconds) from its own 'Simulation elapsed time' / '# of output events' report."""
    exe = os.path.join(ROOT, "oracle", "_ref", "PHOLD")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ, SCALESIM_NUM_THR=str(threads))
    for _ in range(6):      # start-up crash of the reference (stopwatch registry race, stopwatch.hpp:50-56): retry
        try:
            p = subprocess.run([exe, str(n_lp), str(n_init), "0.1", "1.0", "1", "50"], env=env, timeout=timeout,
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        except subprocess.TimeoutExpired:
            return None
        if p.returncode != 0:
            continue
        ms = re.search(r"Simulation elapsed time:\s*(\d+)\s*ms", p.stderr)
        out = re.search(r"# of output events\s*:\s*(\d+)", p.stderr)
        if ms and out:
            return int(out.group(1)), max(int(ms.group(1)), 1) / 1000.0
    return None


def run_port_once(n_lp, n_init):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    s = oracle_lib.Sim.phold(n_lp, n_init, 0.1, 1.0)
    t0 = time.perf_counter()
    n = s.run_timewarp(4, 100, 1)
    dt = time.perf_counter() - t0
    s.close()
    return n, dt


def cpu_run(budget_s):
    """One CPU pass of the reference on the largest sample expected to fit `budget_s` seconds on this host (the rate is
    calibrated with the 10 000-LP sample first). Returns dict(value, kind, cores, sample)."""
    threads = cpu_threads()
    small = SAMPLES[-1]
    r = run_reference_once(small[0], small[1], threads)
    if not r:
        n, dt = run_port_once(small[0], small[1])
        return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
                "sample": f"oracle port (tw_oracle.cc sequential Time Warp emulation, -O2) PHOLD {small[0]} LPs / {small[1]} "
                          f"init events, {n} committed in {dt:.2f} s"}
    base_rate = r[0] / r[1]
    pick = small
    for smp in SAMPLES[:-1]:
        if smp[2] / (base_rate * smp[3]) <= budget_s:
            pick = smp
            break
    if pick is not small:
        r2 = run_reference_once(pick[0], pick[1], threads, timeout=int(budget_s * 4) + 60)
        if r2:
            r = r2
        else:
            pick = small
    return {"value": r[0] / r[1], "unit": UNIT, "cores": threads, "kind": "reference", "seconds": r[1],
            "sample": f"unmodified reference PHOLD {pick[0]} LPs / {pick[1]} init events / 0.1 / 1.0, 1 MPI rank "
                      f"(stand-in Boost.MPI) x {threads} worker threads, -O0, {r[0]} committed in {r[1]:.2f} s "
                      f"(its own 'Simulation elapsed time', includes serial printing)"}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    per_run = 150.0 / max(1, args.steps + args.warmup)
    res = []
    for i in range(args.warmup + args.steps):
        r = cpu_run(per_run)
        if i >= args.warmup:
            res.append(r)
    val = statistics.mean(x["value"] for x in res)
    base = dict(res[-1]); base["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * statistics.mean(x["seconds"] for x in res), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "PHOLD (config 2 shape) bounded sample on host cores: " + base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# Secondary workload: grid TrafficSim (BASELINE.json config 3: 1024 x 1024 torus grid, 10M trips, row-stripe partition)
# ------------------------------------------------------------------------------------------------
B_ALG_TRAFFIC = 392.0   # algorithmic bytes per committed TrafficSim event


def traffic_measure(args, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks, log):
    import numpy as np
    import scalesim_b200 as ssb
    from scalesim_b200 import traffic
    t0 = time.perf_counter()
    sc = traffic.synthetic_grid(args.traffic_grid, args.traffic_grid, args.traffic_trips, seed=1)
    sc.partition = traffic.stripe_partition(args.traffic_grid, args.traffic_grid, world) if world > 1 else None
    log(f"traffic scenario: {sc.trip_id.size} trips, {sc.route_nodes.size} route nodes, built in {time.perf_counter() - t0:.1f} s")
    per = (args.traffic_trips + world - 1) // world
    eng = traffic.make_engine(sc, bucket_ticks=args.traffic_bucket, window_buckets=args.traffic_window, rank=rank, world=world,
                              device=local_rank, max_events=int(per * 2.0) + (1 << 20), max_batch=per + (1 << 16),
                              rounds_per_sync=32, load=False)
    ssb.dist.init_engine_comm(eng, rank, world, p2p=args.exchange == "p2p")
    ev = traffic.pack_events(sc)
    if world > 1:
        ev = ev[sc.partition[ev["dst"]] % world == rank]
    ev = np.ascontiguousarray(ev)
    del sc
    ms, committed, rounds, last = [], 0, 0, None
    for i in range(2 + max(1, min(args.steps, 3))):
        eng.reset()
        eng.load_events(ev)
        barrier()
        st = eng.run()
        barrier()
        if i >= 2:
            ms.append(max_over_ranks(st.loop_ms))
            committed += sum_over_ranks(st.committed)
            rounds += st.rounds
            last = st
    n = len(ms)
    value = committed / (sum(ms) / 1000.0)
    eng.close()
    return {"metric": "committed events/sec, grid TrafficSim", "value": value, "unit": UNIT, "ms_per_step": sum(ms) / n,
            "committed_per_step": committed // n, "rounds_per_step": rounds // n, "rollbacks": last.rollbacks,
            "antis": last.antis_sent, "bucket_ticks": args.traffic_bucket, "window_buckets": args.traffic_window,
            "workload": f"TrafficSim {args.traffic_grid}x{args.traffic_grid} torus grid, {args.traffic_trips} trips "
                        f"(8..64 hops), finish 10800 s, {world} row-stripe partition(s) (BASELINE.json config 3)",
            "whole_path_roofline": {"alg_bytes_per_event": B_ALG_TRAFFIC, "achieved_gbs": value / world * B_ALG_TRAFFIC / 1e9,
                                    "frac": value / world * B_ALG_TRAFFIC / 1e9 / peaks()[0]}}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lps-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--init-per-gpu", type=int, default=16_000_000)
    ap.add_argument("--remote", type=float, default=0.1)
    ap.add_argument("--bucket", type=int, default=10000)
    ap.add_argument("--window", type=int, default=1)
    ap.add_argument("--rounds-per-sync", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) digest check against the CPU oracle")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="cross-partition event exchange")
    ap.add_argument("--no-traffic", action="store_true", help="skip the secondary grid TrafficSim measurement")
    ap.add_argument("--traffic-only", action="store_true")
    ap.add_argument("--traffic-grid", type=int, default=1024)
    ap.add_argument("--traffic-trips", type=int, default=10_000_000)
    ap.add_argument("--traffic-bucket", type=int, default=11)
    ap.add_argument("--traffic-window", type=int, default=1)
    args = ap.parse_args()

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the ONE JSON line and nothing else
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import scalesim_b200 as ssb
    from scalesim_b200 import phold

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")      # plumbing only (unique id, barriers, max over ranks); data path = NCCL in the engine
    torch.cuda.set_device(local_rank)

    verbose = bool(os.environ.get("BENCH_VERBOSE"))

    def log(msg):
        if verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        return ssb.dist.max_over_ranks(x, world)

    def sum_over_ranks(x):
        return ssb.dist.sum_over_ranks(x, world)

    if args.traffic_only:
        t = traffic_measure(args, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks, log)
        if rank == 0:
            print(json.dumps(t), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    n_lp = args.lps_per_gpu * world
    n_init = args.init_per_gpu * world
    tables = phold.build_tables(args.remote, 1.0)
    lat = tables[0]

    # initial events of this rank in pinned host memory (torch = plumbing for the pinned allocation)
    ev_np = phold.init_events(n_lp, n_init, lat, rank, world)
    n_ev = ev_np.size
    # ... in the compact 32-byte layout of the C ABI (ssb_load_events_raw); the widening to ssb_event would double the
    # host-to-device bytes without adding information
    ev_pin = torch.empty(n_ev * ssb.RAW_RECORD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    ev_view = ev_pin.numpy().view(ssb.RAW_RECORD_DTYPE)
    ev_view[:] = ssb.events_to_raw(ev_np)
    del ev_np
    per_rank_committed_cap = int(FULL_COMMITTED_1M * args.lps_per_gpu / 1_000_000 * 1.15) + (1 << 20)
    keep = not args.no_e2e
    eng = phold.make_engine(n_lp, n_init, args.remote, 1.0, tables=tables, bucket_ticks=args.bucket,
                            window_buckets=args.window, rank=rank, world=world, device=local_rank,
                            max_committed=per_rank_committed_cap if keep else 0,
                            max_events=int(args.init_per_gpu * 1.6) + (1 << 20), rounds_per_sync=args.rounds_per_sync,
                            load=False)
    ssb.dist.init_engine_comm(eng, rank, world, p2p=args.exchange == "p2p")
    rec_pin = torch.empty(per_rank_committed_cap * 24, dtype=torch.uint8).pin_memory() if keep else None

    def one_step():
        """device-resident timing: inputs already in HBM when the timed region (the window loop) starts"""
        eng.reset()
        eng.load_events_raw_ptr(ev_pin.data_ptr(), n_ev)
        barrier()
        st = eng.run()                  # loop_ms: CUDA events on the engine's stream around the whole window loop
        barrier()
        return st

    def one_e2e_step():
        barrier()
        t0 = time.perf_counter()
        eng.reset()
        eng.load_events_raw_ptr(ev_pin.data_ptr(), n_ev)                   # H2D from pinned host memory
        # window loop on the device; the committed records of finished rounds stream to pinned host memory meanwhile
        st, got = eng.run_drain_packed_ptr(rec_pin.data_ptr(), per_rank_committed_cap)
        dg = eng.digest()
        barrier()
        return st, dg, got, time.perf_counter() - t0

    log(f"engine ready: {n_ev} initial events on this rank")
    sampler = ClockSampler(local_rank)      # samples from the first warm-up step to the last timed step (same load)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        st = one_step()
        log(f"warm-up step: loop {st.loop_ms:.2f} ms, rounds {st.rounds}, committed {st.committed}, remote {st.remote_sent}")
    step_ms, committed_total, rounds_total, launches_total = [], 0, 0, 0
    last = None
    for _ in range(args.steps):
        st = one_step()
        step_ms.append(max_over_ranks(st.loop_ms))
        committed_total += sum_over_ranks(st.committed)
        rounds_total += st.rounds
        launches_total += launches(st, world)
        last = st
    # parity (untimed): the digest of the union of the partitions' committed records must equal the CPU oracle's
    # closure of the same workload (tests/oracle_lib.py = the checker; SURVEY.md App. A canonical form, sim_obj.hpp:66-77)
    got_digest = ssb.dist.all_gather_digest(eng.digest(), world)
    parity = None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        t0 = time.perf_counter()
        want_digest = oracle_lib.phold_closure_digest(n_lp, n_init, args.remote, 1.0)
        parity = {"checked": True, "equal": tuple(got_digest) == tuple(want_digest), "oracle": "closure (tw_oracle.cc)",
                  "committed": int(got_digest[0]), "oracle_seconds": round(time.perf_counter() - t0, 2)}
        log(f"parity: {parity}")
        if not parity["equal"]:
            print(json.dumps({"error": "committed digest differs from the oracle", "got": list(got_digest),
                              "want": list(want_digest)}), file=sys.stderr, flush=True)
            raise SystemExit(3)
    clocks = sampler.stop() if rank == 0 else None
    total_s = sum(step_ms) / 1000.0
    value = committed_total / total_s
    committed_per_step = committed_total // args.steps

    # per-kernel device time of one more (untimed) step, CUDA events on the engine's stream
    eng.set_flags(ssb.FLAG_KERNEL_TIMING)
    stp = one_step()
    kt = eng.kernel_times()
    eng.set_flags(0)
    peak, peak_src = peaks()
    busy = sum(ms for ms, _ in kt.values())
    dom = max((k for k in kt if k in B_ALG_KERNEL or k == "k_route"), key=lambda k: kt[k][0])
    dom_ms, dom_launches = kt[dom]
    ev_rank = stp.committed
    achieved = B_ALG_KERNEL.get(dom, 0.0) * ev_rank / (dom_ms / 1000.0) / 1e9 if dom_ms > 0 else 0.0
    traffic_bytes, traffic_src = None, None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)
        nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic_bytes, traffic_src = nt[dom]["dram_bytes_per_launch"], nt["_source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic_bytes, "traffic_source": traffic_src,
                "alg_bytes_per_launch": B_ALG_KERNEL.get(dom, 0.0) * ev_rank / max(dom_launches, 1),
                "peak_source": peak_src,
                "alg_bytes_per_event": B_ALG_KERNEL.get(dom), "launches": dom_launches,
                "avg_launch_ms": dom_ms / max(dom_launches, 1),
                "kernel_ms": {k: round(v[0], 4) for k, v in kt.items()},
                "kernel_share": {k: round(v[0] / busy, 4) if busy else 0 for k, v in kt.items()},
                "whole_path": {"alg_bytes_per_event": B_ALG, "achieved": value / world * B_ALG / 1e9,
                               "frac": value / world * B_ALG / 1e9 / peak}}

    # end to end through the host API
    e2e = None
    if keep:
        log("e2e warm-up")
        one_e2e_step()
        log("e2e warm-up done")
        ts, ok = [], True
        want = None
        for _ in range(max(1, min(args.steps, 3))):
            st, dg, got, dt = one_e2e_step()
            ts.append(max_over_ranks(dt))
            ok = ok and got == st.committed
            log(f"e2e step: {dt * 1000:.1f} ms, drained {got}")
        e2e_committed = sum_over_ranks(st.committed)
        e2e = {"value": e2e_committed / statistics.mean(ts), "unit": UNIT,
               # whole-job bytes per step (all ranks): initial events in, committed records + digest + statistics out
               "h2d_bytes_per_step": int(sum_over_ranks(n_ev * ssb.RAW_RECORD_DTYPE.itemsize)),
               "d2h_bytes_per_step": int(sum_over_ranks(st.committed * 24 + 32 + 1296)),
               "ms_per_step": 1000.0 * statistics.mean(ts), "all_records_drained": bool(ok)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_run(30.0)

    eng.close()
    del ev_pin, rec_pin
    traffic_res = None
    if not args.no_traffic:
        traffic_res = traffic_measure(args, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks, log)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": statistics.mean(step_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64 (u32/u64 packed on device)", "data": "synthetic",
            "config": {"workload": f"PHOLD {n_lp} LPs, {n_init} initial events, exponential increment (lambda 1.0), "
                                   f"{args.remote:g} remote, finish 500000 ticks, partition lp % {world} "
                                   f"(BASELINE.json config 2 per GPU)",
                       "bucket_ticks": args.bucket, "window_buckets": args.window,
                       "exchange": ("peer-memory writes over NVLink, GVT min-reduction + barrier through the same mailboxes" if args.exchange == "p2p"
                                    else "NCCL send/recv + allreduce-min") if world > 1 else "none (1 GPU)",
                       "committed_per_step": committed_per_step, "rounds_per_step": rounds_total // args.steps,
                       "l2": "inputs larger than L2 (event pool + tables > 600 MB per GPU); no flush between steps",
                       "rollbacks": last.rollbacks, "antis": last.antis_sent, "irregular_turns": last.fix_turns},
            "parity": parity,
            "gpu_launches": launches_total,
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "traffic": traffic_res,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
