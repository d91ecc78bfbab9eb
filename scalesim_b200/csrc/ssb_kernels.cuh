// ssb_kernels.cuh — the Time Warp core as CUDA kernels (sm_100a).
//
// One round (= one optimistic window step) is, in stream order:
//   k_plan      GVT (block min-reduction over the calendar; k_min + ncclAllReduce(min) first when world > 1),
//               window bounds, gather list, commit list
//   k_commit    stream-compaction of every bucket that fell below GVT into the committed log + digest; the last
//               block to finish does the fossil collection (chunks back to the pool)
//   k_hist      arrivals of the window: per-LP histogram (rank inside the LP's segment)
//   k_segs      per-LP segment allocation (block-aggregated), chain-head note, long-segment list
//   k_scatter   arrivals -> per-LP segments
//   k_sortheavy one warp per LP with a long segment: bitonic sort in shared memory
//   k_exec      one thread per EVENT, in per-LP segment order: place in the LP's key order, handler,
//               processed-chain link, sent-log, outbox (speculative on the LP state)
//   k_fixup(+_heavy) the LPs that are not plain (anti-message, duplicate, straggler, state-mutating handler):
//               annihilate, roll back, re-execute sequentially, emit anti-messages
//   k_route     outbox -> calendar buckets / per-destination-partition send buffers
//               (block-aggregated atomics; anti-messages first, then positives)
//   [exchange + k_route of the receive buffer when world > 1]
// Reference semantics restated per kernel; rule numbers R1..R10 refer to SURVEY.md Appendix A.
#pragma once
#include "ssb_device.cuh"
#include "ssb_models.cuh"

namespace ssb {

constexpr int kRouteThreads = 512;
constexpr int kRouteItems = 8;
constexpr int kRouteTile = kRouteThreads * kRouteItems;

// flag bits of lpflag[lp]
constexpr u32 PL_IRREGULAR = 1u;   // anti-message, duplicate key or possible straggler in the segment
constexpr u32 PL_MUTATED = 2u;     // a handler invocation changed the LP state
constexpr u32 PL_SORTED = 4u;      // the segment is sorted by (key, arrival order)
constexpr u32 PL_FIX = PL_IRREGULAR | PL_MUTATED;
constexpr u32 kScanMax = 96;       // longer segments are sorted by k_sortheavy instead of scanned per event
constexpr u32 kLightMax = 16;      // k_fixup sorts segments up to this length in registers/local memory
constexpr u32 kHeavyCap = 2048;    // entries per warp in the shared-memory bitonic sort
constexpr int kHeavyWarps = 2;

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    u64 w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u64 warp_xor_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// GVT candidate (R7): the smallest key any unprocessed event of this partition can have. After a round every
// integrated event has been executed, so the unprocessed events are exactly the unconsumed tails of the calendar
// buckets; the start time of the first bucket with such a tail is a valid lower bound ("any value <= every LP's
// local time") and gives the same commit frontier as the exact minimum, because commit works on whole
// buckets. Events dropped beyond finish_time count as (finish_time, 0). Warp-shuffle + block min-reduction.
// Replaces scheduler::min_locals (process_scheduler.hpp:83-90) + the local half of mpi_gsync::check_sync
// (global_sync.hpp:95-150).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_local_min(const View& v, u64* sm) {
  Ctl* c = v.ctl;
  u64 m = KEY_MAX;
  for (u32 b = threadIdx.x; b < v.nb; b += blockDim.x)
    if (v.fill[b] != v.consumed[b]) { u64 k = make_key(b * v.bucket_ticks, 0); m = k < m ? k : m; }
  m = warp_min_u64(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : KEY_MAX;
    m = warp_min_u64(m);
    if (threadIdx.x == 0) {
      u64 bm = c->n_beyond ? make_key(v.finish_time, 0) : KEY_MAX;
      m = bm < m ? bm : m;
      // peer-memory exchange: events this partition wrote into other partitions' mailboxes are in flight until the
      // receivers integrate them after the all-reduce; the sender accounts for them (the reference folds the
      // timestamp of every sent message into the sender's local minimum the same way, sender_receiver.hpp:63-67)
      c->local_min = c->sent_min < m ? c->sent_min : m;
    }
  }
}

__global__ void k_min(View v) {
  __shared__ u64 sm[32];
  block_local_min(v, sm);
}

// ------------------------------------------------------------------------------------------------
// k_plan — GVT := (all-reduced) local_min. Finish test of runner.hpp:392. Commit list = buckets entirely below GVT
// (R8/R9). Gather list = unconsumed tails of the buckets inside the optimistic window
// [bucket(GVT), bucket(GVT) + window).
// ------------------------------------------------------------------------------------------------
__global__ void k_plan(View v, int use_global_gvt) {
  __shared__ u64 sm[32];
  Ctl* c = v.ctl;
  if (!use_global_gvt) {     // single partition: the min-reduction is fused here
    block_local_min(v, sm);
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  u64 gvt = use_global_gvt ? c->gvt : c->local_min;
  c->gvt = gvt;
  u32 t = gvt == KEY_MAX ? 0xFFFFFFFFu : key_time(gvt);
  bool done = t >= v.finish_time;
  u32 fin_b = (v.finish_time + v.bucket_ticks - 1) / v.bucket_ticks;   // buckets [0, fin_b) hold times < finish
  if (fin_b > v.nb) fin_b = v.nb;
  u32 cur = done ? fin_b : t / v.bucket_ticks;
  if (cur > fin_b) cur = fin_b;
  c->cur_abs = cur;
  // commit list
  u32 nc = 0, tot = 0;
  for (u32 b = c->commit_next; b < cur; ++b) {
    u32 f = v.fill[b];
    if (f) { v.commit[nc++] = make_uint4(b, f, tot, 0); tot += f; }
  }
  c->commit_next = cur > c->commit_next ? cur : c->commit_next;
  c->n_commit = nc;
  c->commit_total = tot;
  // snapshot ring: rounds whose window lies below GVT are dead (their snapshots can be reused)
  while (c->rq_tail != c->rq_head && c->rq_bound[c->rq_tail % RQ_LEN] <= cur) c->rq_tail++;
  c->snap_tail = c->rq_tail == c->rq_head ? c->snap_head : c->rq_snap[c->rq_tail % RQ_LEN];
  // gather list
  u32 ng = 0, total = 0;
  u32 bound = cur;
  if (!done) {
    bound = cur + v.window < fin_b ? cur + v.window : fin_b;
    for (u32 b = cur; b < bound; ++b) {
      u32 avail = v.fill[b] - v.consumed[b];
      if (!avail) continue;
      u32 take = avail < v.cap_batch - total ? avail : v.cap_batch - total;
      if (take) {
        v.gather[ng++] = make_uint4(b, v.consumed[b], take, total);
        total += take;
        v.consumed[b] += take;
      }
      if (take < avail) { bound = b + 1; break; }   // batch capacity reached: the rest of this bucket stays pending
    }
    if (c->rq_head - c->rq_tail < RQ_LEN) {
      c->rq_bound[c->rq_head % RQ_LEN] = bound;
      c->rq_snap[c->rq_head % RQ_LEN] = c->snap_head;
      c->rq_head++;
    } else {
      c->rq_bound[(c->rq_head - 1) % RQ_LEN] = bound;
    }
  }
  c->bound_abs = bound;
  c->n_gather = ng;
  c->n_arr = total;
  c->n_processed += total;   // k_fixup corrects this for the LPs it re-executes
  c->seg_cursor = 0;
  c->n_anti = 0;
  c->n_extra = 0;
  c->n_heavy = 0;
  c->n_fix = 0;
  c->n_fixheavy = 0;
  c->done = done ? 1u : 0u;
  c->round++;
  for (int i = 0; i < MAX_WORLD; ++i) c->send_count[i] = 0;
  c->sent_min = KEY_MAX;
  if (v.p2p) {
    // the mailbox half that was just integrated (parity of the round that ended) is free again
    const u32 par = (c->round - 1) & 1u;
    for (u32 r = 0; r < v.world; ++r) v.mbox_cnt[(par * v.world + r) * 32] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// k_commit — R8: every event with receive_time < finish_time that survived (not annihilated) is output exactly
// once, when its bucket falls below GVT. Stream compaction into the committed log (warp ballot -> block offsets ->
// one atomic per 256 events), plus the order-independent digest. Replaces eventq::std_out + sim_event::result_out
// (queue.hpp:203-211, sim_obj.hpp:66-77), which print one line per event serially on the main thread.
// The last block to finish does R9, fossil collection: chunks of the committed buckets return to the free stack
// (replaces eventq::release / release_cancel / stateq::release, queue.hpp:159-177, 292-302).
// ------------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(256) k_commit(View v, typename M::Data md) {
  Ctl* c = v.ctl;
  const u32 total = c->commit_total;
  const u32 nc = c->n_commit;
  if (nc == 0) return;
  u64 cnt = 0, hsum = 0, hxor = 0, hsum2 = 0;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ u32 w_cnt[8];
  __shared__ u64 s_pos;
  __shared__ u64 s_dg[8][4];
  __shared__ u32 s_last;
  // every thread of a block iterates the same number of times (block-level compaction inside the loop)
  for (u32 base = blockIdx.x * 256u; base < total; base += gridDim.x * 256u) {
    u32 i = base + threadIdx.x;
    bool keep = false;
    Ev e;
    u32 tg = 0;
    if (i < total) {
      u32 lo = 0, hi = nc;        // locate the bucket: binary search over the commit list offsets
      while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (v.commit[mid].z <= i) lo = mid; else hi = mid;
      }
      uint4 ce = v.commit[lo];
      u32 slot = slot_of(v, ce.x, i - ce.z);
      e = load_ev(v.ev + slot);
      const unsigned char pf = v.pflag[slot];
      keep = !(e.flags & F_CANCEL) && !(pf & P_DEAD) && key_time(e.key) < v.finish_time;
      if (keep) {
        if (!(pf & P_PROCESSED)) atomicOr(&c->error, E_INTERNAL);
        tg = M::tag(md, e);
        u64 h = record_hash(e, tg);
        cnt += 1; hsum += h; hxor ^= h; hsum2 += mix64(h + 0xD1B54A32D192ED03ULL);
      }
    }
    if (v.cap_log) {
      u32 bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) w_cnt[warp] = __popc(bal);
      __syncthreads();
      if (threadIdx.x == 0) {
        u32 tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { u32 a = w_cnt[w]; w_cnt[w] = tot; tot += a; }
        s_pos = tot ? atomicAdd(&c->log_count, (u64)tot) : 0ull;
      }
      __syncthreads();
      if (keep) {
        u64 pos = s_pos + w_cnt[warp] + __popc(bal & ((1u << lane) - 1u));
        if (pos < v.cap_log) { e.flags = tg; store_ev(v.log + pos, e); }
        else atomicAdd(&c->log_dropped, 1ull);
      }
      __syncthreads();
    }
  }
  // digest: warp shuffle -> shared memory -> one set of atomics per block
  cnt = warp_sum_u64(cnt); hsum = warp_sum_u64(hsum); hxor = warp_xor_u64(hxor); hsum2 = warp_sum_u64(hsum2);
  if (lane == 0) { s_dg[warp][0] = cnt; s_dg[warp][1] = hsum; s_dg[warp][2] = hxor; s_dg[warp][3] = hsum2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 a = 0, b = 0, x = 0, d = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += s_dg[w][0]; b += s_dg[w][1]; x ^= s_dg[w][2]; d += s_dg[w][3]; }
    if (a) {
      atomicAdd(&c->digest[0], a);
      atomicAdd(&c->digest[1], b);
      atomicXor(&c->digest[2], x);
      atomicAdd(&c->digest[3], d);
      atomicAdd(&c->n_committed, a);
    }
    __threadfence();
    s_last = atomicAdd(&c->commit_blocks, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  const u32 chs = 1u << v.ch_log;
  for (u32 k = 0; k < nc; ++k) {
    uint4 ce = v.commit[k];
    u32 b = ce.x;
    u32 nch = (ce.y + chs - 1) >> v.ch_log;
    for (u32 j = threadIdx.x; j < nch; j += blockDim.x) {
      u32* entry = &v.chunk_tab[(size_t)b * v.maxc + j];
      u32 ch = *entry;
      *entry = NIL;
      u32 pos = atomicAdd(&c->free_top, 1u);
      v.free_stack[pos] = ch;
    }
    if (threadIdx.x == 0) { v.fill[b] = 0; v.consumed[b] = 0; }
  }
  if (threadIdx.x == 0) c->commit_blocks = 0;
}

// ------------------------------------------------------------------------------------------------
// k_hist — arrivals of this window, in calendar (slot) order. Replaces the inbox append of lp::buffer
// (logical_process.hpp:115-127): rank = position inside the LP's inbox.
// ------------------------------------------------------------------------------------------------
__global__ void k_hist(View v) {
  Ctl* c = v.ctl;
  const u32 n = c->n_arr;
  const u32 ng = c->n_gather;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    u32 g = 0, hi = ng;            // binary search: last gather entry whose offset is <= i
    while (hi - g > 1) {
      u32 mid = (g + hi) >> 1;
      if (v.gather[mid].w <= i) g = mid; else hi = mid;
    }
    uint4 ge = v.gather[g];
    u32 slot = slot_of(v, ge.x, ge.y + (i - ge.w));
    Ev e = load_ev(v.ev + slot);
    u32 lp = part_local(v, e.dst);
    u32 rank = atomicAdd(&v.cnt[lp], 1u);
    // work record = the event with `flags` replaced by its arrival index (bit 31 = anti-message); from here to
    // k_exec the event pool itself is not touched again
    e.flags = i | ((e.flags & F_CANCEL) ? 0x80000000u : 0u);
    store_ev(v.tmp_work + i, e);
    v.tmp_lp[i] = lp;
    v.tmp_rank[i] = rank;
  }
}

// ------------------------------------------------------------------------------------------------
// k_segs — one thread per LP: allocate the LP's segment of this round; 256 LPs share one atomic on the cursor
// (warp scan + block scan). Notes the LP's processed-chain head for this round and lists the long segments.
// Replaces scheduler::queue/dequeue (process_scheduler.hpp:55-81): "which LPs have work in this window".
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_segs(View v) {
  Ctl* c = v.ctl;
  if (c->n_arr == 0) return;
  __shared__ u32 w_seg[8], b_seg;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (u32 base = blockIdx.x * 256u; base < v.n_lp_local; base += gridDim.x * 256u) {
    u32 lp = base + threadIdx.x;
    u32 n = lp < v.n_lp_local ? v.cnt[lp] : 0u;
    u32 incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    if (lane == 31) w_seg[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 ts = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { u32 a = w_seg[w]; w_seg[w] = ts; ts += a; }
      b_seg = ts ? atomicAdd(&c->seg_cursor, ts) : 0u;
    }
    __syncthreads();
    if (n) {
      const u32 start = b_seg + w_seg[warp] + incl - n;
      v.lpw[lp] = make_uint4(start, n, v.last_slot[lp], v.last_time[lp]);
      v.lpflag[lp] = 0;
      v.cnt[lp] = 0;
      if (n > kScanMax) v.heavy[atomicAdd(&c->n_heavy, 1u)] = make_uint4(lp, start, n, 0);
    }
    __syncthreads();
  }
}

// first flagger of an LP puts it on the fix list (light or heavy by segment length)
__device__ __forceinline__ void flag_lp(const View& v, u32 lp, u32 s0, u32 n, u32 bit) {
  u32 old = atomicOr(&v.lpflag[lp], bit);
  if ((old & PL_FIX) == 0) {
    if (n <= kLightMax || (old & PL_SORTED)) v.fix[atomicAdd(&v.ctl->n_fix, 1u)] = make_uint4(lp, s0, n, 0);
    else v.fixheavy[atomicAdd(&v.ctl->n_fixheavy, 1u)] = make_uint4(lp, s0, n, 0);
  }
}

// k_scatter — arrivals -> per-LP segments of work records (arrival order inside a segment).
__global__ void k_scatter(View v) {
  const u32 n = v.ctl->n_arr;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Ev e = load_ev(v.tmp_work + i);
    store_ev(v.work + v.lpw[v.tmp_lp[i]].x + v.tmp_rank[i], e);
  }
}

// calendar slot of the arrival with index i (inverse of k_hist's enumeration)
__device__ __forceinline__ u32 slot_of_arrival(const View& v, u32 i) {
  u32 g = 0, hi = v.ctl->n_gather;
  while (hi - g > 1) {
    u32 mid = (g + hi) >> 1;
    if (v.gather[mid].w <= i) g = mid; else hi = mid;
  }
  const uint4 ge = v.gather[g];
  return slot_of(v, ge.x, ge.y + (i - ge.w));
}

// ------------------------------------------------------------------------------------------------
// lp_turn — one LP's turn in a round, sequentially: what runner::run_lp_aggr does (runner.hpp:530-570) for every
// event of the window instead of switch_lp_interval() of them. The segment is sorted by (key, arrival order) (R1).
//   1. straggler test against the newest processed event; roll back: walk the processed chain, emit the
//      anti-message of every undone event's output, restore the pre-state from the snapshot ring,
//      put the undone events back in front of the merge                  — R5 (queue.hpp:98-105, 286-290)
//   2. merge inbox and undone events in key order; per key apply the inbox items in arrival order:
//      positive = insert unless present, anti = erase if present           — R4/R6 (queue.hpp:82-96)
//   3. execute survivors in key order; record sent-log + snapshot          — R2 (runner.hpp:544-568)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sorted_less(const Sorted& a, const Sorted& b) {
  return a.key < b.key || (a.key == b.key && (a.seq & 0x7FFFFFFFu) < (b.seq & 0x7FFFFFFFu));
}

template <class M>
struct Turn {
  const View& v;
  const typename M::Data& md;
  u32 last_slot, last_time;
  u32 st[M::kStateWords];
  bool st_dirty;
  u32 err;
  u32 n_proc, n_undone, n_anti, n_annih, n_dup;

  __device__ Turn(const View& v_, const typename M::Data& md_) : v(v_), md(md_) {}

  __device__ __forceinline__ void emit_anti(u64 key, u32 dst, u32 src, u32 send) {
    u32 pos = atomicAdd(&v.ctl->n_anti, 1u);
    if (pos >= v.cap_aux) { err |= E_BATCH; return; }
    Ev a;
    a.key = key; a.dst = dst; a.src = src; a.send = send; a.p0 = 0; a.p1 = 0; a.flags = F_CANCEL;
    store_ev(v.anti_box + pos, a);
    ++n_anti;
  }

  // run the handler on the event in `slot`; out_idx = outbox position for its output
  __device__ __forceinline__ void process(u32 slot, u32 out_idx) {
    Ev e = load_ev(v.ev + slot);
    u32 pre[M::kStateWords];
#pragma unroll
    for (int w = 0; w < M::kStateWords; ++w) pre[w] = st[w];
    Ev out;
    bool has = M::handle(md, e, st, out, err);
    Meta m;
    m.link = last_slot;
    m.prevtime = last_time;
    m.sent_dst = NIL;
    m.snap = NIL;
    m.sent_key = 0;
    if (has) {
      bool changed = false;
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) changed |= (pre[w] != st[w]);
      if (changed) {
        // state saving: the pre-state goes to the snapshot ring (copy-on-change)
        u32 s = atomicAdd(&v.ctl->snap_head, 1u);
        if (s - v.ctl->snap_tail >= v.cap_snap) err |= E_SNAP;
        u32 si = s % v.cap_snap;
#pragma unroll
        for (int w = 0; w < M::kStateWords; ++w) v.snap[(size_t)si * M::kStateWords + w] = pre[w];
        m.snap = si;
        st_dirty = true;
      }
      m.sent_dst = out.dst;
      m.sent_key = out.key;
      store_ev(v.outbox + out_idx, out);
    } else {
      // no event produced: nothing recorded, state update dropped (runner.hpp:552-553, quirk Q2)
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) st[w] = pre[w];
      v.outbox[out_idx].key = KEY_MAX;
    }
    store_meta(v.meta + slot, m);
    v.pflag[slot] = P_PROCESSED;
    last_slot = slot;
    last_time = key_time(e.key);
    ++n_proc;
  }

  __device__ __forceinline__ u32 extra_idx() {
    u32 pos = atomicAdd(&v.ctl->n_extra, 1u);
    if (pos >= v.cap_aux) { err |= E_BATCH; pos = v.cap_aux - 1; }
    return v.cap_batch + pos;
  }

  // old_slot/old_time = the LP's chain head before this round. The LP's arrivals own the outbox entries
  // [s0, s0+n) — one each — so whatever k_exec may already have written there for this LP is overwritten here.
  __device__ void run(u32 lp, u32 s0, u32 n, u32 old_slot, u32 old_time) {
    err = 0; n_proc = n_undone = n_anti = n_annih = n_dup = 0;
    st_dirty = false;
    const Sorted* seg = v.sorted + s0;
    last_slot = old_slot;
    last_time = old_time;
#pragma unroll
    for (int w = 0; w < M::kStateWords; ++w) st[w] = v.state[(size_t)lp * M::kStateWords + w];

    // 1. rollback
    u32 U = NIL;   // undone events, ascending key order, linked through Meta::link
    const u64 m = seg[0].key;
    const u32 mtime = key_time(m);
    if (last_slot != NIL && last_time >= mtime) {
      u32 cur = last_slot, curtime = last_time;
      bool any = false;
      while (cur != NIL && curtime >= mtime) {
        Ev e = load_ev(v.ev + cur);
        if (e.key < m) break;
        Meta mt = load_meta(v.meta + cur);
        if (mt.sent_dst != NIL) emit_anti(mt.sent_key, mt.sent_dst, e.dst, key_time(e.key));
        if (mt.snap != NIL) {
#pragma unroll
          for (int w = 0; w < M::kStateWords; ++w) st[w] = v.snap[(size_t)mt.snap * M::kStateWords + w];
          st_dirty = true;
        }
        v.pflag[cur] = 0;
        v.meta[cur].link = U;
        U = cur;
        cur = mt.link;
        curtime = mt.prevtime;
        ++n_undone;
        any = true;
      }
      last_slot = cur;
      last_time = curtime;
      if (any) atomicAdd(&v.ctl->n_rollbacks, 1ull);
    }

    // 2 + 3. merge and execute
    u32 ia = 0;
    while (ia < n || U != NIL) {
      u64 ka = ia < n ? seg[ia].key : KEY_MAX;
      u64 ku = KEY_MAX;
      if (U != NIL) ku = v.ev[U].key;
      u64 k = ka < ku ? ka : ku;
      u32 present = NIL;      // slot of the event currently in the queue under key k
      u32 out_idx = NIL;
      if (ku == k) {
        present = U;
        U = v.meta[present].link;
      }
      while (ia < n && seg[ia].key == k) {
        const Sorted a = seg[ia];
        const u32 my_out = s0 + ia;
        if (a.seq & 0x80000000u) {
          v.pflag[a.slot] = P_DEAD;                     // the anti-message itself is consumed
          v.outbox[my_out].key = KEY_MAX;
          if (present != NIL) {                         // erase (queue.hpp:87-90)
            v.pflag[present] = P_DEAD;
            present = NIL;
            if (out_idx != NIL) { v.outbox[out_idx].key = KEY_MAX; out_idx = NIL; }
            ++n_annih;
          }
        } else if (present == NIL) {
          present = a.slot;                             // insert (queue.hpp:92)
          out_idx = my_out;
        } else {
          v.pflag[a.slot] = P_DEAD;                     // duplicate key: second insert ignored
          v.outbox[my_out].key = KEY_MAX;
          ++n_dup;
        }
        ++ia;
      }
      if (present != NIL) {
        // a re-executed (rolled back) event has no inbox position: its output goes to the extra region
        process(present, out_idx != NIL ? out_idx : extra_idx());
      }
    }

    v.last_slot[lp] = last_slot;
    v.last_time[lp] = last_time;
    if (st_dirty) {
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) v.state[(size_t)lp * M::kStateWords + w] = st[w];
    }
  }
};

// ------------------------------------------------------------------------------------------------
// warp-cooperative sort of one segment by (key, arrival order): bitonic network in shared memory (n <= kHeavyCap),
// else one lane sorts in place. Returns true (warp-uniform) if the segment holds an anti-message or a duplicate key.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Sorted sorted_from_work(const View& v, const Ev* wk) {
  Sorted x;
  x.key = wk->key;
  x.seq = wk->flags;
  x.slot = slot_of_arrival(v, x.seq & 0x7FFFFFFFu);
  return x;
}

// builds sorted[s0 .. s0+n) = (key, slot, arrival) of work[s0 .. s0+n), sorted
__device__ __forceinline__ bool warp_sort_segment(const View& v, u32 s0, u32 n, Sorted* buf, u32 lane) {
  u32 flag = 0;
  Sorted* seg = v.sorted + s0;
  if (n <= kHeavyCap) {
    u32 P = 1;
    while (P < n) P <<= 1;
    for (u32 i = lane; i < P; i += 32) {
      Sorted x;
      if (i < n) x = sorted_from_work(v, v.work + s0 + i); else { x.key = KEY_MAX; x.slot = NIL; x.seq = 0x7FFFFFFFu; }
      buf[i] = x;
    }
    __syncwarp();
    for (u32 k2 = 2; k2 <= P; k2 <<= 1) {
      for (u32 j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        for (u32 i = lane; i < P; i += 32) {
          u32 ixj = i ^ j2;
          if (ixj > i) {
            Sorted x = buf[i], y = buf[ixj];
            bool asc = (i & k2) == 0;
            if (sorted_less(y, x) == asc) { buf[i] = y; buf[ixj] = x; }
          }
        }
        __syncwarp();
      }
    }
    for (u32 i = lane; i < n; i += 32) {
      Sorted x = buf[i];
      seg[i] = x;
      flag |= x.seq >> 31;
      if (i && buf[i - 1].key == x.key) flag = 1;
    }
  } else {
    for (u32 i = lane; i < n; i += 32) seg[i] = sorted_from_work(v, v.work + s0 + i);
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
      for (u32 i = 1; i < n; ++i) {
        Sorted x = seg[i];
        u32 j = i;
        while (j > 0) {
          Sorted y = seg[j - 1];
          if (!sorted_less(x, y)) break;
          seg[j] = y;
          --j;
        }
        if (j != i) seg[j] = x;
      }
    }
    __threadfence_block();
    __syncwarp();
    for (u32 i = lane; i < n; i += 32) {
      Sorted x = seg[i];
      flag |= x.seq >> 31;
      if (i && seg[i - 1].key == x.key) flag = 1;
    }
  }
  __syncwarp();
  return __any_sync(0xffffffffu, flag != 0);
}

// ------------------------------------------------------------------------------------------------
// k_sortheavy — one warp per LP whose segment is too long for the per-event scan of k_exec (> kScanMax, the hot
// LPs): sort it (R1) and classify it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHeavyWarps * 32) k_sortheavy(View v) {
  extern __shared__ uint4 smem_heavy[];
  Ctl* c = v.ctl;
  const u32 nh = c->n_heavy;
  if (nh == 0) return;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Sorted* buf = reinterpret_cast<Sorted*>(smem_heavy) + (size_t)warp * kHeavyCap;
  for (u32 h = blockIdx.x * kHeavyWarps + warp; h < nh; h += gridDim.x * kHeavyWarps) {
    uint4 a = v.heavy[h];
    const u32 lp = a.x, s0 = a.y, n = a.z;
    const Sorted* seg = v.sorted + s0;
    bool irregular = warp_sort_segment(v, s0, n, buf, lane);
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
      const uint4 w = v.lpw[lp];
      if (w.z != NIL && key_time(seg[0].key) <= w.w) irregular = true;     // possible straggler
      atomicOr(&v.lpflag[lp], PL_SORTED);
      if (irregular) flag_lp(v, lp, s0, n, PL_IRREGULAR);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// k_exec — one thread per EVENT of the window, in per-LP segment order over the work records (a coalesced stream;
// the threads of one segment share its cache lines): find the event's place in its LP's key order (R1), run the
// handler on the LP's current state, link the event into the LP's processed chain, record the sent-log entry, write
// the produced event to the outbox. Skew-free: a hot LP's events are spread over many threads.
//
// Place in key order: segments up to kScanMax events are scanned by each of their events (all threads of a segment
// read the same few cache lines): predecessor = largest smaller key, "last" = no larger key. The scan also finds the
// irregular turns — an anti-message, a duplicate key, or a key not later than the LP's newest processed event
// (straggler, R5) — which are left to the sequential lp_turn of k_fixup. Longer segments were sorted by k_sortheavy
// and are searched.
//
// This is the speculative form of runner::run_lp_aggr (runner.hpp:530-570): all events of an LP are executed against
// the same state. That is exactly the sequential result as long as no invocation changes the state (neither shipped
// application ever does). The moment one does, the LP is flagged and k_fixup re-executes its whole segment
// sequentially — handlers that mutate state keep exact semantics.
// ------------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(256) k_exec(View v, typename M::Data md) {
  Ctl* c = v.ctl;
  const u32 n_arr = c->n_arr;
  u32 err = 0;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n_arr; j += gridDim.x * blockDim.x) {
    Ev e = load_ev(v.work + j);
    const u32 seq = e.flags;
    const u32 lp = part_local(v, e.dst);
    const uint4 w = v.lpw[lp];
    const u32 s0 = w.x, n = w.y;
    const u32 fl = v.lpflag[lp];
    if (fl & PL_FIX) continue;                          // handled by k_fixup
    u32 pred_slot = w.z, pred_time = w.w;
    bool is_last = true;
    bool bad = (seq >> 31) != 0;
    if (w.z != NIL && key_time(e.key) <= w.w) bad = true;
    if (fl & PL_SORTED) {
      const Sorted* seg = v.sorted + s0;
      u32 lo = 0, hi = n;                               // position of my key in the sorted segment
      while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (seg[mid].key <= e.key) lo = mid; else hi = mid;
      }
      if (lo > 0) { const Sorted sp = seg[lo - 1]; pred_slot = sp.slot; pred_time = key_time(sp.key); }
      is_last = lo == n - 1;
    } else if (n > 1) {
      u64 pk = 0;
      u32 pseq = 0;
      bool have = false;
      const Ev* seg = v.work + s0;
      for (u32 k = 0; k < n; ++k) {
        if (s0 + k == j) continue;
        const u64 tk = seg[k].key;
        if (seg[k].flags >> 31) bad = true;
        if (tk == e.key) bad = true;
        else if (tk < e.key) {
          if (!have || tk > pk) { pk = tk; pseq = seg[k].flags; have = true; }
        } else {
          is_last = false;
        }
      }
      if (have && !bad) { pred_slot = slot_of_arrival(v, pseq & 0x7FFFFFFFu); pred_time = key_time(pk); }
    }
    if (bad) { flag_lp(v, lp, s0, n, PL_IRREGULAR); continue; }
    const u32 slot = slot_of_arrival(v, seq);
    e.flags = 0;
    u32 st[M::kStateWords], pre[M::kStateWords];
#pragma unroll
    for (int q = 0; q < M::kStateWords; ++q) { st[q] = v.state[(size_t)lp * M::kStateWords + q]; pre[q] = st[q]; }
    Ev out;
    bool has = M::handle(md, e, st, out, err);
    bool changed = false;
#pragma unroll
    for (int q = 0; q < M::kStateWords; ++q) changed |= (pre[q] != st[q]);
    if (has && changed) { flag_lp(v, lp, s0, n, PL_MUTATED); continue; }   // state-mutating handler
    Meta m;
    m.link = pred_slot;
    m.prevtime = pred_time;
    m.sent_dst = NIL;
    m.snap = NIL;
    m.sent_key = 0;
    if (has) {
      m.sent_dst = out.dst;
      m.sent_key = out.key;
      store_ev(v.outbox + j, out);
    } else {
      v.outbox[j].key = KEY_MAX;
    }
    store_meta(v.meta + slot, m);
    v.pflag[slot] = P_PROCESSED;
    if (is_last) { v.last_slot[lp] = slot; v.last_time[lp] = key_time(e.key); }
  }
  if (err) atomicOr(&c->error, err);
}

// ------------------------------------------------------------------------------------------------
// k_fixup / k_fixup_heavy — the LPs whose turn is not plain (anti-message, duplicate, straggler, state-mutating
// handler): sort the segment if it is not sorted yet, then the exact sequential lp_turn. k_fixup: one thread per LP
// (short or already sorted segments); k_fixup_heavy: one warp per LP (bitonic sort, then lane 0 runs the turn).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fixup_stats(Ctl* c, u32 lane, u64 n_proc, u64 n_redo, u64 n_undone, u64 n_anti,
                                            u64 n_annih, u64 n_dup, u32 err) {
  n_proc = warp_sum_u64(n_proc);
  n_redo = warp_sum_u64(n_redo);
  n_undone = warp_sum_u64(n_undone);
  n_anti = warp_sum_u64(n_anti);
  n_annih = warp_sum_u64(n_annih);
  n_dup = warp_sum_u64(n_dup);
  if (lane == 0) {
    if (n_proc != n_redo) atomicAdd(&c->n_processed, n_proc - n_redo);   // k_plan counted every arrival once
    if (n_undone) atomicAdd(&c->n_undone, n_undone);
    if (n_anti) atomicAdd(&c->n_antis, n_anti);
    if (n_annih) atomicAdd(&c->n_annihilated, n_annih);
    if (n_dup) atomicAdd(&c->n_dups, n_dup);
  }
  if (err) atomicOr(&c->error, err);
}

template <class M>
__global__ void k_fixup(View v, typename M::Data md) {
  Ctl* c = v.ctl;
  const u32 nf = c->n_fix;
  if (nf == 0) return;
  u64 n_proc = 0, n_undone = 0, n_anti = 0, n_annih = 0, n_dup = 0, n_redo = 0;
  u32 err = 0;
  const u32 lane = threadIdx.x & 31;
  const u32 nthreads = gridDim.x * blockDim.x;
  for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < nf; base += nthreads) {
    u32 k = base + lane;
    if (k < nf) {
      const uint4 f = v.fix[k];
      const u32 lp = f.x, s0 = f.y, n = f.z;
      const uint4 w = v.lpw[lp];
      if (!(v.lpflag[lp] & PL_SORTED)) {
        Sorted* seg = v.sorted + s0;
        Sorted loc[kLightMax];
        for (u32 q = 0; q < n; ++q) loc[q] = sorted_from_work(v, v.work + s0 + q);
        for (u32 q = 1; q < n; ++q) {
          Sorted x = loc[q];
          u32 j = q;
          while (j > 0 && sorted_less(x, loc[j - 1])) { loc[j] = loc[j - 1]; --j; }
          loc[j] = x;
        }
        for (u32 q = 0; q < n; ++q) seg[q] = loc[q];
      }
      n_redo += n;
      Turn<M> t(v, md);
      t.run(lp, s0, n, w.z, w.w);
      n_proc += t.n_proc; n_undone += t.n_undone; n_anti += t.n_anti; n_annih += t.n_annih; n_dup += t.n_dup;
      err |= t.err;
    }
  }
  __syncwarp();
  fixup_stats(c, lane, n_proc, n_redo, n_undone, n_anti, n_annih, n_dup, err);
}

template <class M>
__global__ void __launch_bounds__(kHeavyWarps * 32) k_fixup_heavy(View v, typename M::Data md) {
  extern __shared__ uint4 smem_heavy[];
  Ctl* c = v.ctl;
  const u32 nf = c->n_fixheavy;
  if (nf == 0) return;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Sorted* buf = reinterpret_cast<Sorted*>(smem_heavy) + (size_t)warp * kHeavyCap;
  u64 n_proc = 0, n_undone = 0, n_anti = 0, n_annih = 0, n_dup = 0, n_redo = 0;
  u32 err = 0;
  for (u32 h = blockIdx.x * kHeavyWarps + warp; h < nf; h += gridDim.x * kHeavyWarps) {
    const uint4 f = v.fixheavy[h];
    const u32 lp = f.x, s0 = f.y, n = f.z;
    warp_sort_segment(v, s0, n, buf, lane);
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
      const uint4 w = v.lpw[lp];
      n_redo += n;
      Turn<M> t(v, md);
      t.run(lp, s0, n, w.z, w.w);
      n_proc += t.n_proc; n_undone += t.n_undone; n_anti += t.n_anti; n_annih += t.n_annih; n_dup += t.n_dup;
      err |= t.err;
    }
    __syncwarp();
  }
  __syncwarp();
  fixup_stats(c, lane, n_proc, n_redo, n_undone, n_anti, n_annih, n_dup, err);
}

// ------------------------------------------------------------------------------------------------
// k_route — emit: outbox -> calendar buckets (local destination) or per-partition send buffers.
// Replaces event_com::send_event (event_communicator.hpp:90-103) + lp::buffer. Per tile of 4096 events the block
// builds a shared-memory histogram over targets and claims one range per target with a single global atomic
// (block-aggregated). The appended event is at once part of the bucket's unconsumed tail, which is what the GVT
// reduction looks at (R3: the destination's local time drops at once). Events at or beyond finish_time are never
// executed nor committed (R8), so they are counted and dropped (GVT treats them as (finish_time, 0)).
//   which: 0 = anti box, 1 = main + extra outbox, 2 = receive buffer, 3 = host-loaded staging (outbox[0..load_count))
//   filter: 0 = all, 1 = anti-messages only, 2 = positives only (the receive side keeps antis ahead of positives)
// ------------------------------------------------------------------------------------------------
// pops a chunk off the free stack (chunk 0 is a sacrificial chunk: it keeps everyone moving after an overflow)
__device__ __forceinline__ u32 alloc_chunk(const View& v) {
  Ctl* c = v.ctl;
  u32 top = atomicSub(&c->free_top, 1u);
  if (top == 0 || top > 0x7FFFFFFFu) {
    atomicAdd(&c->free_top, 1u);
    atomicOr(&c->error, E_POOL);
    return 0;
  }
  atomicMin(&c->free_low, top - 1);
  return v.free_stack[top - 1];
}
// chunk id published by whoever owns the chunk's first position (bounded wait: a logic error cannot hang the GPU)
__device__ __forceinline__ u32 wait_chunk(const View& v, const u32* entry) {
  u32 ch = __ldcg(entry);
  for (u32 spin = 0; ch == NIL && spin < (1u << 22); ++spin) ch = *reinterpret_cast<const volatile u32*>(entry);
  if (ch == NIL) atomicOr(&v.ctl->error, E_INTERNAL);
  return ch;
}

__global__ void __launch_bounds__(kRouteThreads) k_route(View v, int which, u32 peer, u32 filter) {
  extern __shared__ u64 smem64[];
  Ctl* c = v.ctl;
  const Ev* src;
  u32 n0, n1 = 0;        // main count, extra count (extra region starts at cap_batch)
  if (which == 0) { src = v.anti_box; n0 = c->n_anti < v.cap_aux ? c->n_anti : v.cap_aux; }
  else if (which == 1) { src = v.outbox; n0 = c->n_arr; n1 = c->n_extra < v.cap_aux ? c->n_extra : v.cap_aux; }
  else if (which == 2) {
    if (v.p2p) {
      const u32 par = c->round & 1u;
      src = v.mbox_data + ((size_t)par * v.world + peer) * v.cap_send;
      n0 = v.mbox_cnt[(par * v.world + peer) * 32];
      if (n0 > v.cap_send) { n0 = v.cap_send; if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&c->error, E_SENDBUF); }
    } else {
      src = v.recvbuf + (size_t)peer * v.cap_send; n0 = c->recv_count[peer];
    }
  }
  else { src = v.outbox; n0 = c->load_count; }
  const u32 total = n0 + n1;
  if (total == 0) return;
  // shared-memory targets: [0, rs) = buckets base .. base+rs-1, rs = "beyond", rs+1+r = partition r.
  // Buckets further ahead than the table (tiny bucket_ticks) take one global atomic per event instead.
  const u32 rs = v.route_slots;
  const u32 nt = rs + 1 + v.world;
  const u32 base = c->commit_next;
  u32* s_cnt = reinterpret_cast<u32*>(smem64);          // [nt] events of the tile per target
  u32* s_base = s_cnt + nt;                             // [nt] first position claimed in the target
  u32* s_ch0 = s_base + nt;                             // [nt] chunk of that first position
  u32* s_ch1 = s_ch0 + nt;                              // [nt] the next chunk, if the range crosses into it
  const u32 chm = (1u << v.ch_log) - 1u;
  const u32 T_NONE = NIL, T_DIRECT = NIL - 1u;
  __shared__ u64 s_sent_min;
  if (threadIdx.x == 0) s_sent_min = KEY_MAX;
  __syncthreads();
  for (u32 tile = blockIdx.x * kRouteTile; tile < total; tile += gridDim.x * kRouteTile) {
    for (u32 t = threadIdx.x; t < nt; t += kRouteThreads) s_cnt[t] = 0;
    __syncthreads();
    u32 tgt[kRouteItems], rnk[kRouteItems];
    // pass 1: target of every event (first half of the record is enough), rank inside the tile's share of the target
#pragma unroll
    for (int it = 0; it < kRouteItems; ++it) {
      u32 i = tile + it * kRouteThreads + threadIdx.x;
      tgt[it] = T_NONE;
      rnk[it] = 0;
      if (i < total) {
        u32 idx = i < n0 ? i : v.cap_batch + (i - n0);
        const uint4 h = reinterpret_cast<const uint4*>(src + idx)[0];     // key lo, key hi, dst, src
        const u64 key = ((u64)h.y << 32) | h.x;
        bool take = key != KEY_MAX;
        if (take && filter) {
          const u32 fl = src[idx].flags;
          take = filter == 1u ? (fl & F_CANCEL) != 0 : (fl & F_CANCEL) == 0;
        }
        if (take) {
          const u32 dst = h.z;
          if (dst >= v.n_lp) { atomicOr(&c->error, E_RANGE); continue; }
          const u32 tm = key_time(key);
          const u32 owner = v.world > 1 ? part_owner(v, dst) : v.rank;
          u32 t;
          if (owner != v.rank) {
            if (which == 3) { atomicOr(&c->error, E_RANGE); continue; }   // loaded events must be owned by this rank
            t = rs + 1 + owner;
          } else if (tm >= v.finish_time) {
            t = rs;
          } else {
            const u32 b = tm / v.bucket_ticks;
            if (b < base) { atomicOr(&c->error, E_INTERNAL); continue; }   // causality: nothing may arrive below GVT
            t = b - base < rs ? b - base : T_DIRECT;
            if (t == T_DIRECT) rnk[it] = atomicAdd(&v.fill[b], 1u);        // final position
          }
          tgt[it] = t;
          if (t != T_DIRECT) rnk[it] = atomicAdd(&s_cnt[t], 1u);
        }
      }
    }
    __syncthreads();
    // one thread per target: claim the range, allocate + publish the chunks that start inside it. No thread waits
    // for anything before every publish of the block is out (next barrier), so the waits below cannot form a cycle.
    for (u32 t = threadIdx.x; t < nt; t += kRouteThreads) {
      u32 n = s_cnt[t];
      if (!n) continue;
      if (t < rs) {
        const u32 b = base + t;
        const u32 p0 = atomicAdd(&v.fill[b], n);
        s_base[t] = p0;
        const u32 j0 = p0 >> v.ch_log, j1 = (p0 + n - 1) >> v.ch_log;
        u32 ch0 = NIL, ch1 = NIL;
        for (u32 j = (p0 & chm) ? j0 + 1 : j0; j <= j1; ++j) {
          u32 ch = alloc_chunk(v);
          if (j >= v.maxc) { atomicOr(&c->error, E_POOL); continue; }
          atomicExch(&v.chunk_tab[(size_t)b * v.maxc + j], ch);      // publish
          if (j == j0) ch0 = ch;
          if (j == j0 + 1) ch1 = ch;
        }
        s_ch0[t] = ch0;
        s_ch1[t] = ch1;
      } else if (t == rs) {
        atomicAdd(&c->n_beyond, (u64)n);
      } else {
        const u32 r = t - rs - 1;
        if (v.p2p) s_base[t] = atomicAdd_system(&v.peer_cnt[r][((c->round & 1u) * v.world + v.rank) * 32], n);
        else s_base[t] = atomicAdd(&c->send_count[r], n);
        atomicAdd(&c->n_remote, (u64)n);
      }
    }
    // far buckets (beyond the table): the owner of a chunk's first position allocates it
#pragma unroll
    for (int it = 0; it < kRouteItems; ++it) {
      if (tgt[it] == T_DIRECT && (rnk[it] & chm) == 0) {
        u32 i = tile + it * kRouteThreads + threadIdx.x;
        u32 idx = i < n0 ? i : v.cap_batch + (i - n0);
        const u32 b = key_time(src[idx].key) / v.bucket_ticks;
        u32 ch = alloc_chunk(v);
        if ((rnk[it] >> v.ch_log) >= v.maxc) atomicOr(&c->error, E_POOL);
        else atomicExch(&v.chunk_tab[(size_t)b * v.maxc + (rnk[it] >> v.ch_log)], ch);
      }
    }
    __syncthreads();
    // the chunk a range starts in was opened by another block unless the range starts at the chunk's first position
    for (u32 t = threadIdx.x; t < rs; t += kRouteThreads) {
      if (s_cnt[t] && s_ch0[t] == NIL) {
        const u32 j0 = s_base[t] >> v.ch_log;
        if (j0 < v.maxc) s_ch0[t] = wait_chunk(v, &v.chunk_tab[(size_t)(base + t) * v.maxc + j0]);
      }
    }
    __syncthreads();
    // copy the records (re-read from the outbox: L1/L2 hits) to their slots
#pragma unroll
    for (int it = 0; it < kRouteItems; ++it) {
      const u32 t = tgt[it];
      if (t == T_NONE || t == rs) continue;
      u32 i = tile + it * kRouteThreads + threadIdx.x;
      u32 idx = i < n0 ? i : v.cap_batch + (i - n0);
      Ev e = load_ev(src + idx);
      e.flags &= F_CANCEL;
      if (t < rs || t == T_DIRECT) {
        u32 p, ch;
        if (t < rs) {
          p = s_base[t] + rnk[it];
          const u32 j = p >> v.ch_log, j0 = s_base[t] >> v.ch_log;
          if (j >= v.maxc) continue;
          ch = j == j0 ? s_ch0[t] : (j == j0 + 1 ? s_ch1[t] : wait_chunk(v, &v.chunk_tab[(size_t)(base + t) * v.maxc + j]));
        } else {
          p = rnk[it];
          if ((p >> v.ch_log) >= v.maxc) continue;
          ch = wait_chunk(v, &v.chunk_tab[(size_t)(key_time(e.key) / v.bucket_ticks) * v.maxc + (p >> v.ch_log)]);
        }
        if (ch == NIL) continue;
        const size_t slot = ((size_t)ch << v.ch_log) | (p & chm);
        store_ev(v.ev + slot, e);
        v.pflag[slot] = 0;
      } else {
        const u32 r = t - rs - 1;
        const u32 p = s_base[t] + rnk[it];
        if (p >= v.cap_send) { atomicOr(&c->error, E_SENDBUF); continue; }
        if (v.p2p) {
          // straight into the destination GPU's mailbox over NVLink
          store_ev(v.peer_data[r] + (((size_t)(c->round & 1u) * v.world + v.rank) * v.cap_send + p), e);
          const u32 tm = key_time(e.key);
          const u64 k = make_key(tm >= v.finish_time ? v.finish_time : tm / v.bucket_ticks * v.bucket_ticks, 0);
          if (k < s_sent_min) atomicMin(&s_sent_min, k);
        } else {
          store_ev(v.sendbuf + (size_t)r * v.cap_send + p, e);
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && s_sent_min != KEY_MAX) atomicMin(&c->sent_min, s_sent_min);
}

// ------------------------------------------------------------------------------------------------
// k_pack — host events (ssb_event, 64 B) -> device records in the outbox staging area, with range checks.
// ------------------------------------------------------------------------------------------------
struct HostEvent { long long id, src, dst, recv, send, p0, p1, flags; };

__global__ void k_pack(View v, const HostEvent* in, u32 n) {
  Ctl* c = v.ctl;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    HostEvent h = in[i];
    bool ok = h.id >= 0 && h.id < 0xFFFFFFFFll && h.dst >= 0 && h.dst < (long long)v.n_lp &&
              h.recv >= 0 && h.recv < 0x80000000ll && h.send >= 0 && h.send < 0x100000000ll &&
              h.src >= -1 && h.src < 0xFFFFFFFFll && h.p0 >= 0 && h.p0 < 0x100000000ll &&
              h.p1 >= 0 && h.p1 < 0x100000000ll;
    Ev e;
    if (!ok) {
      atomicOr(&c->error, E_RANGE);
      e.key = KEY_MAX; e.dst = 0; e.src = 0; e.send = 0; e.p0 = 0; e.p1 = 0; e.flags = 0;
    } else {
      e.key = make_key((u32)h.recv, (u32)h.id);
      e.dst = (u32)h.dst;
      e.src = h.src < 0 ? NIL : (u32)h.src;
      e.send = (u32)h.send;
      e.p0 = (u32)h.p0; e.p1 = (u32)h.p1;
      e.flags = (h.flags & 1) ? F_CANCEL : 0u;
    }
    store_ev(v.outbox + i, e);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) c->load_count = n;
}

// committed log record (device Ev with flags = tag) -> ssb_record (6 x int64)
__global__ void k_unpack_log(const Ev* log, u64 first, u32 n, long long* out) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Ev e = load_ev(log + first + i);
    long long* o = out + (size_t)i * 6;
    o[0] = key_id(e.key);
    o[1] = e.src == NIL ? -1ll : (long long)e.src;
    o[2] = e.send;
    o[3] = e.dst;
    o[4] = key_time(e.key);
    o[5] = e.flags;
  }
}

__global__ void k_fill_u32(u32* p, u32 val, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = val;
}
__global__ void k_iota_u32(u32* p, u32 start, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = start + (u32)i;
}
// PHOLD table packing: latency << 1 | remote
__global__ void k_pack_phold_table(const long long* lat, const int* rem, u32* out, u32 n, Ctl* c) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long l = lat[i];
    if (l < 0 || l >= 0x7FFFFFFFll) { atomicOr(&c->error, E_RANGE); l = 0; }
    out[i] = ((u32)l << 1) | (rem[i] == 1 ? 1u : 0u);
  }
}
__global__ void k_narrow_u32(const long long* in, u32* out, size_t n, Ctl* c) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    long long x = in[i];
    if (x < 0 || x >= 0xFFFFFFFFll) { atomicOr(&c->error, E_RANGE); x = 0; }
    out[i] = (u32)x;
  }
}

}  // namespace ssb
