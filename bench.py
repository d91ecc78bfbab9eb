#!/usr/bin/env python
"""bench.py — committed events/s of the Time Warp hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete simulation through the C ABI.
  N = 1 : BASELINE.json config 2 — PHOLD 1M LPs, 16M initial events, 10 % remote, lambda 1.0, finish 500 000 ticks
          (72 823 220 committed events).
  N > 1 : BASELINE.json config 4 — PHOLD 16M LPs, 256M initial events, 50 % remote, partition lp % N
          (1 165 396 484 committed events; total work fixed: strong scaling). The weak-scaled config 2 (1M LPs per GPU)
          is kept as the extra key "weak_config2".
  value : committed events / s of the window loop with the initial events already in HBM (device time, CUDA events,
          max over ranks) — the reference's "Simulation elapsed time" excludes init the same way (runner.hpp:71-73)
  e2e   : the same metric through the public host API with HOST buffers: H2D of the initial events from
          pinned memory (ssb_load_events_raw, 32 B per event) + window loop + digest + D2H of every committed record
          into pinned memory (ssb_run_drain_packed: 24 B per record, overlapped with the loop), wall clock
  parity: after the timed steps the digest of the union of the partitions' committed records is compared with the
          CPU oracle's closure of the same workload (tests/oracle_lib.py — the checker, never the thing measured);
          a mismatch ends the run with a non-zero exit status.
  optimistic : the same workload with a window of 2 buckets (> lookahead): stragglers, rollbacks, anti-messages —
          same committed digest, its own timing.
  traffic : grid TrafficSim (BASELINE.json config 3), digest checked against the oracle's closure.
  --impl reference : the reference's own CPU implementation (oracle/_ref, unmodified sources; else the oracle port)
          on ONE bounded PHOLD sample (100 000 LPs), same metric/unit.
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

METRIC = "committed events/sec, PHOLD"
UNIT = "events/s"
B_ALG = 380.0   # algorithmic bytes per committed PHOLD event (SURVEY.md §8d / DESIGN.md §5)
# per-kernel split of the 380 B: execute (pop read 48 + table gathers 12 + state read 8 + snapshot write 24 +
# sent-log write 24 + emit write 48), route (bucket read 48 + pending write 48), commit + fossil (read 48 + 40-byte
# result record + fossil key reads 32)
B_ALG_KERNEL = {"k_exec": 164.0, "k_route": 96.0, "k_commit": 120.0}
CONFIG2 = dict(n_lp=1_000_000, n_init=16_000_000, remote=0.1, committed=72_823_220, name="BASELINE.json config 2")
CONFIG4 = dict(n_lp=16_000_000, n_init=256_000_000, remote=0.5, committed=1_165_396_484, name="BASELINE.json config 4")


# kernels per round inside the window-loop graph: plan, commit, exec, route, post = 5; N > 1 adds k_gvt and the
# receive-side route once per global synchronisation; a round with an irregular LP turn (ssb_stats.fix_rounds) adds the
# IF node's six kernels: k_fix_count, k_fix_segs, k_fix_scatter, k_fix_sort, k_fix_turn, k_route(anti).
def launches(st, world):
    return int(st.rounds * 5 + (st.exchanges * 2 if world > 1 else 0) + st.fix_rounds * 6)


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
# CPU side: the reference itself (oracle/_ref) or the oracle port, on ONE bounded sample
# ------------------------------------------------------------------------------------------------
# the bounded PHOLD sample of the config-2 shape (16 events per LP, 0.1 remote): 100 000 LPs, 7 275 345 committed
# events, about 35-40 s of the unmodified reference on 8 worker threads. Both the reference arm and the cpu_baseline
# leg use this one size.
CPU_SAMPLE = (100000, 1600000, 7275345)
CPU_SAMPLE_SMALL = (10000, 160000, 728183)      # fallback when the large one does not finish


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


def cpu_run():
    """One CPU pass of the reference on the bounded sample. Returns dict(value, kind, cores, sample)."""
    threads = cpu_threads()
    for smp in (CPU_SAMPLE, CPU_SAMPLE_SMALL):
        r = run_reference_once(smp[0], smp[1], threads, timeout=400)
        if r:
            return {"value": r[0] / r[1], "unit": UNIT, "cores": threads, "kind": "reference", "seconds": r[1],
                    "sample": f"unmodified reference PHOLD {smp[0]} LPs / {smp[1]} init events / 0.1 / 1.0, 1 MPI rank "
                              f"(stand-in Boost.MPI) x {threads} worker threads, -O0, {r[0]} committed in {r[1]:.2f} s "
                              f"(its own 'Simulation elapsed time', includes serial printing)"}
    smp = CPU_SAMPLE_SMALL
    n, dt = run_port_once(smp[0], smp[1])
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
            "sample": f"oracle port (tw_oracle.cc sequential Time Warp emulation, -O2) PHOLD {smp[0]} LPs / {smp[1]} "
                      f"init events, {n} committed in {dt:.2f} s"}


def reference_arm(args, rank, world):
    """ONE timed pass of the reference on the bounded sample (a pass takes ~40 s: the --steps / --warmup of the GPU arm
    would take half an hour here, so they are reported but not multiplied out)."""
    if rank != 0:
        return
    r = cpu_run()
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": 1000.0 * r["seconds"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "PHOLD (config 2 shape) bounded sample on host cores: " + r["sample"]},
            "cpu_baseline": r,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Env:
    """per-process plumbing: rank / world, barriers and reductions over ranks (torch.distributed gloo)"""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.verbose = bool(os.environ.get("BENCH_VERBOSE"))

    def log(self, msg):
        if self.verbose:
            print(f"[bench rank {self.rank}] {msg}", file=sys.stderr, flush=True)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def max(self, x):
        import scalesim_b200 as ssb
        return ssb.dist.max_over_ranks(x, self.world)

    def sum(self, x):
        import scalesim_b200 as ssb
        return ssb.dist.sum_over_ranks(x, self.world)

    def bcast(self, obj):
        if self.world <= 1:
            return obj
        import torch.distributed as dist
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]


# ------------------------------------------------------------------------------------------------
# PHOLD workload
# ------------------------------------------------------------------------------------------------
def phold_rank_events(n_lp, n_init, lat, rank, world):
    """initial events owned by this rank, in the compact 32-byte layout, in pinned host memory"""
    import numpy as np
    import torch
    import ctypes as C
    import scalesim_b200 as ssb
    if world > 1 and n_lp % world == 0:
        # event id -> LP id % n_lp -> partition lp % world: with n_lp a multiple of world, rank r owns ids = r (mod world)
        lib = ssb.load_library()
        cap = (n_init + world - 1) // world + 1
        out = np.empty(cap, dtype=ssb.EVENT_DTYPE)
        n = lib.ssb_phold_init_events(n_lp, n_init, C.c_void_p(lat.ctypes.data), lat.size, rank, world,
                                      C.c_void_p(out.ctypes.data), out.size)
        if n < 0:
            raise RuntimeError("ssb_phold_init_events failed")
        ev_np = out[:n]
    else:
        from scalesim_b200 import phold
        ev_np = phold.init_events(n_lp, n_init, lat, rank, world)
    n_ev = ev_np.size
    ev_pin = torch.empty(n_ev * ssb.RAW_RECORD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    ev_pin.numpy().view(ssb.RAW_RECORD_DTYPE)[:] = ssb.events_to_raw(ev_np)
    return ev_pin, n_ev


def phold_measure(env, wl, tables, *, window, steps, warmup, want_digest, e2e_steps=0, kernel_timing=False,
                  sample_clocks=False):
    """Runs one PHOLD workload `steps` times; returns the measurement dict (every rank checks the parity verdict)."""
    import scalesim_b200 as ssb
    from scalesim_b200 import phold
    import torch
    args, rank, world = env.args, env.rank, env.world
    n_lp, n_init, remote = wl["n_lp"], wl["n_init"], wl["remote"]
    ev_pin, n_ev = phold_rank_events(n_lp, n_init, tables[0], rank, world)
    per_init = (n_init + world - 1) // world
    keep = e2e_steps > 0
    cap_committed = int(wl["committed"] / world * 1.25) + (1 << 20)
    eng = phold.make_engine(n_lp, n_init, remote, 1.0, tables=tables, bucket_ticks=args.bucket, window_buckets=window,
                            rank=rank, world=world, device=env.local_rank, max_committed=cap_committed if keep else 0,
                            max_events=int(per_init * 1.6) + (1 << 20), rounds_per_sync=args.rounds_per_sync, load=False)
    ssb.dist.init_engine_comm(eng, rank, world, p2p=args.exchange == "p2p")
    env.log(f"engine ready: {n_ev} initial events on this rank ({wl['name']}, window {window})")

    def one_step():
        """device-resident timing: inputs already in HBM when the timed region (the window loop) starts"""
        env.barrier()
        eng.reset()
        env.barrier()                    # every partition has reset (mailboxes) before anyone loads or runs
        eng.load_events_raw_ptr(ev_pin.data_ptr(), n_ev)
        env.barrier()
        st = eng.run()                  # loop_ms: CUDA events on the engine's stream around the whole window loop
        env.barrier()
        return st

    sampler = ClockSampler(env.local_rank) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        st = one_step()
        env.log(f"warm-up step: loop {st.loop_ms:.2f} ms, rounds {st.rounds}, committed {st.committed}, remote {st.remote_sent}, "
                f"rollbacks {st.rollbacks}")
    step_ms, committed_total, rounds_total, launches_total, last = [], 0, 0, 0, None
    for _ in range(steps):
        st = one_step()
        step_ms.append(env.max(st.loop_ms))
        committed_total += env.sum(st.committed)
        rounds_total += st.rounds
        launches_total += launches(st, world)
        last = st
    clocks = sampler.stop() if sampler else None
    # parity (untimed): digest of the union of the partitions' committed records vs the oracle's closure
    got = ssb.dist.all_gather_digest(eng.digest(), world)
    parity = None
    if want_digest is not None:
        parity = {"checked": True, "equal": tuple(got) == tuple(want_digest), "oracle": "closure (oracle/tw_oracle.cc)",
                  "committed": int(got[0])}
        if not parity["equal"]:
            if rank == 0:
                print(json.dumps({"error": "committed digest differs from the oracle", "workload": wl["name"], "window": window,
                                  "got": list(got), "want": list(want_digest)}), file=sys.stderr, flush=True)
            raise SystemExit(3)
    res = {"ms_per_step": statistics.mean(step_ms), "value": committed_total / (sum(step_ms) / 1000.0),
           "committed_per_step": committed_total // steps, "rounds_per_step": rounds_total // steps,
           "launches": launches_total, "rollbacks": env.sum(last.rollbacks), "antis": env.sum(last.antis_sent),
           "undone": env.sum(last.undone), "irregular_turns": env.sum(last.fix_turns), "remote_sent": env.sum(last.remote_sent),
           "parity": parity, "clocks": clocks, "window_buckets": window}

    if kernel_timing:
        # per-kernel device time of one more (untimed) step, CUDA events on the engine's stream (stream path)
        eng.set_flags(ssb.FLAG_KERNEL_TIMING)
        stp = one_step()
        kt = eng.kernel_times()
        eng.set_flags(0)
        peak, peak_src = peaks()
        busy = sum(ms for ms, _ in kt.values())
        dom = max((k for k in kt if k in B_ALG_KERNEL), key=lambda k: kt[k][0])
        dom_ms = kt[dom][0]
        work_launches = stp.rounds - 1          # the last round only finds GVT >= finish_time: its launches are empty
        achieved = B_ALG_KERNEL[dom] * stp.committed / (dom_ms / 1000.0) / 1e9 if dom_ms > 0 else 0.0
        traffic_bytes, traffic_src = None, None
        try:      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)
            nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic_bytes, traffic_src = nt[dom]["dram_bytes_per_launch"], nt["_source"]
        except Exception:
            pass
        res["roofline"] = {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic_bytes, "traffic_source": traffic_src,
            "alg_bytes_per_launch": B_ALG_KERNEL[dom] * stp.committed / max(work_launches, 1),
            "peak_source": peak_src, "alg_bytes_per_event": B_ALG_KERNEL[dom], "launches": work_launches,
            "avg_launch_ms": dom_ms / max(work_launches, 1),
            "kernel_ms": {k: round(v[0], 4) for k, v in kt.items()},
            "kernel_share": {k: round(v[0] / busy, 4) if busy else 0 for k, v in kt.items()},
            "whole_path": {"alg_bytes_per_event": B_ALG, "achieved": res["value"] / world * B_ALG / 1e9,
                           "frac": res["value"] / world * B_ALG / 1e9 / peak}}

    if keep:
        rec_pin = torch.empty(cap_committed * 24, dtype=torch.uint8).pin_memory()

        def one_e2e_step():
            env.barrier()
            eng.reset()
            env.barrier()
            t0 = time.perf_counter()
            eng.load_events_raw_ptr(ev_pin.data_ptr(), n_ev)                   # H2D from pinned host memory
            # window loop on the device; the committed records of finished rounds stream to pinned host memory meanwhile
            st, got_n = eng.run_drain_packed_ptr(rec_pin.data_ptr(), cap_committed)
            dg = eng.digest()
            env.barrier()
            return st, dg, got_n, time.perf_counter() - t0

        one_e2e_step()      # warm-up
        ts, ok = [], True
        for _ in range(e2e_steps):
            st, dg, got_n, dt = one_e2e_step()
            ts.append(env.max(dt))
            ok = ok and got_n == st.committed
            env.log(f"e2e step: {dt * 1000:.1f} ms, drained {got_n}")
        e2e_digest = ssb.dist.all_gather_digest(dg, world)
        if want_digest is not None and tuple(e2e_digest) != tuple(want_digest):
            raise SystemExit(3)
        res["e2e"] = {"value": env.sum(st.committed) / statistics.mean(ts), "unit": UNIT,
                      # whole-job bytes per step (all ranks): initial events in, committed records + digest + statistics out
                      "h2d_bytes_per_step": int(env.sum(n_ev * ssb.RAW_RECORD_DTYPE.itemsize)),
                      "d2h_bytes_per_step": int(env.sum(st.committed * 24 + 32 + 1312)),
                      "ms_per_step": 1000.0 * statistics.mean(ts), "all_records_drained": bool(ok),
                      "note": "H2D load of the step's events + window loop + D2H of every committed record + digest, wall clock, max over ranks"}
        del rec_pin
    eng.close()
    del ev_pin
    return res


def phold_oracle_digest(env, wl):
    """closure digest of a PHOLD workload from the CPU oracle (rank 0 computes, everyone gets it); None = unavailable"""
    d = None
    if env.rank == 0 and not env.args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        try:
            import oracle_lib
            t0 = time.perf_counter()
            d = tuple(int(x) for x in oracle_lib.phold_closure_digest(wl["n_lp"], wl["n_init"], wl["remote"], 1.0,
                                                                      threads=max(4, cpu_threads())))
            env.log(f"oracle closure of {wl['name']}: {d[0]} events in {time.perf_counter() - t0:.1f} s")
        except MemoryError:
            d = None
    return env.bcast(d)


# ------------------------------------------------------------------------------------------------
# Secondary workload: grid TrafficSim (BASELINE.json config 3: 1024 x 1024 torus grid, 10M trips, row-stripe partition)
# ------------------------------------------------------------------------------------------------
B_ALG_TRAFFIC = 392.0   # algorithmic bytes per committed TrafficSim event


def traffic_measure(env):
    import numpy as np
    import scalesim_b200 as ssb
    from scalesim_b200 import traffic
    args, rank, world = env.args, env.rank, env.world
    t0 = time.perf_counter()
    sc = traffic.synthetic_grid(args.traffic_grid, args.traffic_grid, args.traffic_trips, seed=1)
    sc.partition = traffic.stripe_partition(args.traffic_grid, args.traffic_grid, world) if world > 1 else None
    env.log(f"traffic scenario: {sc.trip_id.size} trips, {sc.route_nodes.size} route nodes, built in {time.perf_counter() - t0:.1f} s")
    want = None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        t0 = time.perf_counter()
        s = oracle_lib.Sim.traffic_scenario(sc)
        s.run_closure(max(4, cpu_threads()), keep_records=False)
        want = tuple(int(x) for x in s.digest())
        s.close()
        env.log(f"oracle closure of the traffic scenario: {want[0]} events in {time.perf_counter() - t0:.1f} s")
    want = env.bcast(want)
    per = (args.traffic_trips + world - 1) // world
    eng = traffic.make_engine(sc, bucket_ticks=args.traffic_bucket, window_buckets=args.traffic_window, rank=rank, world=world,
                              device=env.local_rank, max_events=int(per * 2.0) + (1 << 20), max_batch=per + (1 << 16),
                              rounds_per_sync=32, load=False)
    ssb.dist.init_engine_comm(eng, rank, world, p2p=args.exchange == "p2p")
    if world > 1 and args.exchange == "p2p":
        eng.set_sync_interval(args.traffic_sync)
    ev = traffic.pack_events(sc)
    if world > 1:
        ev = ev[sc.partition[ev["dst"]] % world == rank]
    ev = np.ascontiguousarray(ev)
    del sc
    ms, committed, rounds, last = [], 0, 0, None
    for i in range(2 + max(1, min(args.steps, 3))):
        env.barrier()
        eng.reset()
        env.barrier()
        eng.load_events(ev)
        env.barrier()
        st = eng.run()
        env.barrier()
        if i >= 2:
            ms.append(env.max(st.loop_ms))
            committed += env.sum(st.committed)
            rounds += st.rounds
            last = st
    got = ssb.dist.all_gather_digest(eng.digest(), world)
    parity = None
    if want is not None:
        parity = {"checked": True, "equal": tuple(got) == tuple(want), "oracle": "closure (oracle/tw_oracle.cc)", "committed": int(got[0])}
        if not parity["equal"]:
            if rank == 0:
                print(json.dumps({"error": "TrafficSim digest differs from the oracle", "got": list(got), "want": list(want)}),
                      file=sys.stderr, flush=True)
            raise SystemExit(3)
    n = len(ms)
    value = committed / (sum(ms) / 1000.0)
    # per-kernel device time of one more (untimed) step on the stream path, CUDA events around every launch
    eng.set_flags(ssb.FLAG_KERNEL_TIMING)
    env.barrier()
    eng.reset()
    env.barrier()
    eng.load_events(ev)
    env.barrier()
    eng.run()
    env.barrier()
    kt = eng.kernel_times()
    eng.set_flags(0)
    eng.close()
    return {"metric": "committed events/sec, grid TrafficSim", "value": value, "unit": UNIT, "ms_per_step": sum(ms) / n,
            "committed_per_step": committed // n, "rounds_per_step": rounds // n, "rollbacks": last.rollbacks,
            "antis": last.antis_sent, "bucket_ticks": args.traffic_bucket, "window_buckets": args.traffic_window,
            "windows_per_synchronisation": args.traffic_sync if world > 1 else 1,
            "parity": parity, "scaling": "strong",
            "kernel_ms": {k: round(v[0], 4) for k, v in kt.items()}, "kernel_launches": {k: int(v[1]) for k, v in kt.items()},
            "workload": f"TrafficSim {args.traffic_grid}x{args.traffic_grid} torus grid, {args.traffic_trips} trips "
                        f"(8..64 hops), finish 10800 s, {world} row-stripe partition(s) (BASELINE.json config 3)",
            "whole_path_roofline": {"alg_bytes_per_event": B_ALG_TRAFFIC, "achieved_gbs": value / world * B_ALG_TRAFFIC / 1e9,
                                    "frac": value / world * B_ALG_TRAFFIC / 1e9 / peaks()[0]}}


# ------------------------------------------------------------------------------------------------
# Exact-differential simulation at scale (BASELINE.json config 5): --diff_init on the config-3 grid with the history
# streaming out while the loop runs, then --diff_repeat with added trips (AE queries)
# ------------------------------------------------------------------------------------------------
def exact_diff_measure(env):
    import numpy as np
    import scalesim_b200 as ssb
    from scalesim_b200 import traffic
    from scalesim_b200.store import HistoryStore
    args = env.args
    g, n_trips, n_add = args.traffic_grid, args.diff_trips, args.diff_added
    sc = traffic.synthetic_grid(g, g, n_trips, seed=1)
    extra = traffic.synthetic_grid(g, g, n_add, seed=9, max_departure=6000)
    extra_ids = 100_000_000 + np.arange(n_add, dtype=np.int64)
    n_nodes = int(sc.route_nodes.size)
    # --diff_init: every committed event + its sent-log entry leaves through the pinned staging buffers while the
    # graph loop runs (ssb_run_stream_history); the sink appends to the host arrays the repeat run loads from
    eng = traffic.make_engine(sc, bucket_ticks=args.traffic_bucket, history=True, max_events=2 * n_trips + (1 << 20),
                              max_batch=n_trips + (1 << 16), max_committed=n_nodes + 1024)
    # destination of the sink: pinned host memory (page-locked, already faulted in), as a writer thread's queue would be
    import torch
    rec_t = torch.empty(n_nodes * ssb.RAW_RECORD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    sent_t = torch.empty(n_nodes * ssb.SENT_RECORD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    rec = rec_t.numpy().view(ssb.RAW_RECORD_DTYPE)
    sent = sent_t.numpy().view(ssb.SENT_RECORD_DTYPE)
    fill = [0]

    rec_b, sent_b = rec_t.numpy(), sent_t.numpy()      # byte views: structured-dtype assignment is not a plain memcpy in numpy

    def sink(r, s):
        k = fill[0]
        rec_b[k * 32:(k + r.size) * 32] = r.view(np.uint8)
        sent_b[k * 16:(k + s.size) * 16] = s.view(np.uint8)
        fill[0] = k + r.size

    t0 = time.perf_counter()
    st = eng.run_stream_history(sink)
    init_wall = time.perf_counter() - t0
    init_digest = eng.digest()
    eng.close()
    n_hist = fill[0]
    hs = HistoryStore(rec[:n_hist], sent[:n_hist], None, sc.route_nodes)
    # --diff_repeat
    t0 = time.perf_counter()
    eng = traffic.make_repeat_engine(sc, hs, (extra_ids, extra.trip_dep, extra.route_start, extra.route_nodes),
                                     bucket_ticks=args.traffic_bucket, keep_records=True)
    load_wall = time.perf_counter() - t0
    t0 = time.perf_counter()
    st2 = eng.run()
    rep = eng.drain_committed_packed()
    repeat_wall = time.perf_counter() - t0
    eng.close()
    # exactness: vehicles do not interact, so ONE run over trips + added trips = init + the added trips' lines; the
    # rest of the repeat output is recorded history that had to be re-executed (it must reproduce recorded lines)
    new = rep[rep["id"] >= 100_000_000]
    redone = rep[rep["id"] < 100_000_000]
    parity = None
    if not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        full = traffic.Scenario(sc.n_lp, sc.state_ids, sc.road_start, sc.road_dst, sc.road_speed, sc.road_len,
                                np.concatenate([sc.trip_id, extra_ids]), np.concatenate([sc.trip_dep, extra.trip_dep]),
                                np.concatenate([sc.route_start, sc.route_start[-1] + extra.route_start[1:]]),
                                np.concatenate([sc.route_nodes, extra.route_nodes]))
        o = oracle_lib.Sim.traffic_scenario(full)
        o.run_closure(max(4, cpu_threads()), keep_records=False)
        want = tuple(int(x) for x in o.digest())
        o.close()
        got = ssb.dist.combine_digests([init_digest, ssb.digest_records(ssb.packed_to_records(new))])
        # re-executed history: every such record is one of the recorded ones (same key, destination, source, send time)
        hk = np.sort(rec[:n_hist]["key"])
        rk = (redone["recv_time"].astype(np.uint64) << np.uint64(32)) | redone["id"].astype(np.uint64)
        pos = np.searchsorted(hk, rk)
        pos[pos >= hk.size] = hk.size - 1
        redone_ok = bool(np.all(hk[pos] == rk)) if redone.size else True
        parity = {"checked": True, "equal": tuple(got) == want and redone_ok, "rule": "digest(init) + digest(new lines) == oracle closure of trips + added trips; re-executed records are recorded ones"}
        if not parity["equal"]:
            print(json.dumps({"error": "exact-differential run is not exact", "got": list(got), "want": list(want), "redone_ok": redone_ok}),
                  file=sys.stderr, flush=True)
            raise SystemExit(3)
    return {"workload": f"TrafficSim {g}x{g} torus grid, {n_trips} trips; --diff_init, then --diff_repeat with {n_add} added trips "
                        f"(BASELINE.json config 5)",
            "init": {"committed": st.committed, "loop_ms": st.loop_ms, "wall_ms": 1000 * init_wall,
                     "history_bytes": int(n_hist) * 48, "events_per_s": st.committed / init_wall,
                     "note": "wall = window loop + D2H of every (event, sent-log entry) pair through 2 pinned staging buffers + host sink"},
            "repeat": {"load_history_wall_ms": 1000 * load_wall, "run_and_drain_wall_ms": 1000 * repeat_wall, "loop_ms": st2.loop_ms,
                       "rounds": st2.rounds, "committed": st2.committed, "new_lines": int(new.size), "reexecuted_history": int(redone.size),
                       "rollbacks": st2.rollbacks, "antis": st2.antis_sent,
                       "fraction_of_full_run_reexecuted": st2.committed / max(1, st.committed + int(new.size))},
            "parity": parity}


# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# Input at scale: CSV in the reference's format -> native readers -> scenario image -> HBM, host wall
# clock; the engine started from the image must commit what the engine fed through the array calls commits. A side
# measurement: a failure is reported in the key, it does not take the bench line down.
# ------------------------------------------------------------------------------------------------
def ingest_measure(env):
    try:
        from scalesim_b200 import scenario
        res = scenario.ingest_measure(env.args.traffic_grid, env.args.traffic_grid, env.args.ingest_trips,
                                      bucket_ticks=env.args.traffic_bucket, device=env.local_rank)
        if not (res.get("digest_equal") and res.get("read_equals_generator")):
            res["error"] = "the engine started from the image and the array path disagree"
        return res
    except Exception as e:      # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bucket", type=int, default=10000)
    ap.add_argument("--window", type=int, default=1)
    ap.add_argument("--opt-window", type=int, default=2, help="window (in buckets) of the optimistic sub-line")
    ap.add_argument("--rounds-per-sync", type=int, default=8)
    ap.add_argument("--workload", default="auto", choices=["auto", "config2", "config4"],
                    help="auto: config 2 on one GPU, config 4 on several")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) digest checks against the CPU oracle")
    ap.add_argument("--no-optimistic", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaled config 2 extra line")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="cross-partition event exchange")
    ap.add_argument("--no-traffic", action="store_true", help="skip the secondary grid TrafficSim measurement")
    ap.add_argument("--traffic-only", action="store_true")
    ap.add_argument("--exact-diff-only", action="store_true")
    ap.add_argument("--traffic-grid", type=int, default=1024)
    ap.add_argument("--traffic-trips", type=int, default=10_000_000)
    ap.add_argument("--traffic-bucket", type=int, default=11)
    ap.add_argument("--traffic-window", type=int, default=1)
    ap.add_argument("--traffic-sync", type=int, default=2, help="N > 1: local windows per global synchronisation (TrafficSim)")
    ap.add_argument("--no-exact-diff", action="store_true", help="skip the exact-differential (config 5) measurement")
    ap.add_argument("--diff-trips", type=int, default=2_000_000)
    ap.add_argument("--diff-added", type=int, default=1000)
    ap.add_argument("--no-ingest", action="store_true", help="skip the scenario-ingestion measurement (CSV -> image -> HBM)")
    ap.add_argument("--ingest-only", action="store_true")
    ap.add_argument("--ingest-trips", type=int, default=500_000)
    args = ap.parse_args()

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the ONE JSON line and nothing else
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from scalesim_b200 import phold

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # plumbing only (IPC handles, barriers, reductions over ranks); the data path is inside the engine
        dist.init_process_group("gloo", timeout=datetime.timedelta(minutes=60))
    env = Env(args)
    torch.cuda.set_device(env.local_rank)

    if args.traffic_only:
        t = traffic_measure(env)
        if rank == 0:
            print(json.dumps(t), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    if args.exact_diff_only:
        print(json.dumps(exact_diff_measure(env)), flush=True)
        return

    if args.ingest_only:
        print(json.dumps(ingest_measure(env)), flush=True)
        return

    which = args.workload if args.workload != "auto" else ("config2" if world == 1 else "config4")
    wl = dict(CONFIG2 if which == "config2" else CONFIG4)
    if which == "config2" and world > 1:      # explicit request: weak-scale it (1M LPs per GPU)
        wl = dict(n_lp=CONFIG2["n_lp"] * world, n_init=CONFIG2["n_init"] * world, remote=0.1,
                  committed=CONFIG2["committed"] * world, name=f"config 2 x {world} (weak-scaled)")
    tables = phold.build_tables(wl["remote"], 1.0)
    want = phold_oracle_digest(env, wl)
    main_res = phold_measure(env, wl, tables, window=args.window, steps=args.steps, warmup=args.warmup, want_digest=want,
                             e2e_steps=0 if args.no_e2e else max(1, min(args.steps, 3)), kernel_timing=True, sample_clocks=True)
    optimistic = None
    if not args.no_optimistic:
        o = phold_measure(env, wl, tables, window=args.opt_window, steps=max(1, min(args.steps, 3)), warmup=1, want_digest=want)
        optimistic = {k: o[k] for k in ("value", "ms_per_step", "window_buckets", "rounds_per_step", "rollbacks", "undone", "antis",
                                        "irregular_turns", "committed_per_step", "parity")}
        optimistic["unit"] = UNIT
    weak = None
    if world > 1 and which == "config4" and not args.no_weak:
        wl2 = dict(n_lp=CONFIG2["n_lp"] * world, n_init=CONFIG2["n_init"] * world, remote=0.1,
                   committed=CONFIG2["committed"] * world, name=f"config 2 x {world} (weak-scaled: 1M LPs per GPU)")
        tables2 = phold.build_tables(0.1, 1.0)
        want2 = phold_oracle_digest(env, wl2)
        w = phold_measure(env, wl2, tables2, window=1, steps=max(1, min(args.steps, 5)), warmup=2, want_digest=want2, kernel_timing=True)
        weak = {k: w[k] for k in ("value", "ms_per_step", "rounds_per_step", "committed_per_step", "parity", "remote_sent")}
        weak.update(unit=UNIT, scaling="weak", workload=wl2["name"], kernel_ms=w["roofline"]["kernel_ms"],
                    whole_path=w["roofline"]["whole_path"])
        del tables2
    del tables

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_run()

    traffic_res = None
    if not args.no_traffic:
        traffic_res = traffic_measure(env)

    exact_diff = None
    if world == 1 and not args.no_exact_diff and not args.no_traffic:
        exact_diff = exact_diff_measure(env)

    ingest = None
    if rank == 0 and world == 1 and not args.no_ingest and not args.no_traffic:
        ingest = ingest_measure(env)

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if which == "config2" else "strong", "vs_baseline": None,
            "dtype": "int64 (u32/u64 packed on device)", "data": "synthetic",
            "config": {"workload": f"PHOLD {wl['n_lp']} LPs, {wl['n_init']} initial events, exponential increment (lambda 1.0), "
                                   f"{wl['remote']:g} remote, finish 500000 ticks, partition lp % {world} ({wl['name']})",
                       "bucket_ticks": args.bucket, "window_buckets": args.window,
                       "exchange": ("peer-memory writes over NVLink (hand-written kernels), GVT min-reduction + barrier through "
                                    "the same mailboxes; no NCCL call on the data path" if args.exchange == "p2p"
                                    else "NCCL send/recv + allreduce-min") if world > 1 else "none (1 GPU)",
                       "committed_per_step": main_res["committed_per_step"], "rounds_per_step": main_res["rounds_per_step"],
                       "l2": "inputs larger than L2 (event pool + tables > 600 MB per GPU); no flush between steps",
                       "rollbacks": main_res["rollbacks"], "antis": main_res["antis"], "irregular_turns": main_res["irregular_turns"],
                       "remote_sent": main_res["remote_sent"]},
            "parity_checked": bool(main_res["parity"] and main_res["parity"]["equal"]), "parity": main_res["parity"],
            "gpu_launches": main_res["launches"],
            "clocks": main_res["clocks"], "roofline": main_res.get("roofline"), "e2e": main_res.get("e2e"),
            "cpu_baseline": cpu, "optimistic": optimistic, "weak_config2": weak, "traffic": traffic_res, "exact_diff": exact_diff,
            "ingest": ingest,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
