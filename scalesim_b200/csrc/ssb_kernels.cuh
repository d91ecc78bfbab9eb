// ssb_kernels.cuh — the Time Warp core as CUDA kernels (sm_100a).
//
// One round (= one optimistic window step), as the body of the window-loop graph (ssb_engine.cu: build_graph):
//   [k_gvt      N > 1: GVT min-reduction + barrier through the peers' mailboxes; then k_route of the received
//               anti-messages and of the received positives]
//   k_plan      GVT (block min-reduction over the calendar when N = 1), window bounds, gather list, commit list,
//               loop condition of the graph
//   k_commit    (parallel branch) stream-compaction of every bucket that fell below GVT into the committed log +
//               digest (+ sent-log entries for the history store); the last block does the fossil collection
//   k_link      arrivals of the window, in calendar order: count per LP, chain every arrival to the previous
//               arrival of its LP (atomic exchange on the LP's list head), note the chain head, list the long LPs
//   k_segs_heavy  segment allocation for the LPs with more than light_max arrivals
//   k_exec      one thread per EVENT, in calendar order (coalesced): handler, sent-log, outbox (speculative on the LP
//               state); an event of a short LP finds its place in the LP's key order by walking the arrival list, an
//               event of a long LP drops (key, slot, arrival) into the LP's segment
//   [k_sortheavy  IF a segment longer than kRegSortSmall exists: one block per such LP, hybrid bitonic sort + links]
//   k_exec_heavy  one warp per long LP: bitonic sort in registers, processed-chain links from the sorted neighbours
//   [k_fixup + k_route(anti)  IF some LP is not plain (anti-message, duplicate, straggler, state-mutating handler):
//               annihilate, roll back, re-execute sequentially, re-queue, emit anti-messages]
//   k_route     outbox -> calendar buckets / the destination GPU's mailbox (TMA-staged tiles, block-aggregated atomics)
// The same kernels in replay mode (View::replay) rebuild the processed chains of a recorded history (--diff_repeat).
// Reference semantics restated per kernel; rule numbers R1..R13 refer to SURVEY.md Appendix A.
#pragma once
#include "ssb_device.cuh"
#include "ssb_models.cuh"

namespace ssb {

constexpr int kRouteThreads = 256;
constexpr int kRouteItems = 4;
constexpr int kRouteTile = kRouteThreads * kRouteItems;   // records staged in shared memory per tile (32 KB)

// flag bits of an LP in this round (lph[lp].x >> kCntBits)
constexpr u32 PL_IRREGULAR = 1u;   // anti-message, duplicate key or possible straggler in the segment
constexpr u32 PL_MUTATED = 2u;     // a handler invocation changed the LP state
constexpr u32 PL_SORTED = 4u;      // the segment is sorted by (key, arrival order)
constexpr u32 PL_FIX = PL_IRREGULAR | PL_MUTATED;
constexpr u32 kCntBits = 28;       // lph[lp].x = arrivals of the round (bits 0..27) | PL_* << 28
constexpr u32 kCntMask = (1u << kCntBits) - 1u;
constexpr u32 NIL31 = 0x7FFFFFFFu; // end of an LP's arrival list
constexpr u32 kLightMax = 16;      // upper bound of View::light_max: LPs with more arrivals in a round get a segment
constexpr u32 kRegSortSmall = 128; // register sort with up to 4 entries per lane (the every-round kernel)

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
__device__ __forceinline__ void block_local_min(const View& v, Ctl* c, u64* sm) {
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
  block_local_min(v, v.ctl, sm);
}

// ------------------------------------------------------------------------------------------------
// k_gvt — peer-memory GVT (replaces mpi_gsync::check_sync, global_sync.hpp:95-150, and the NCCL all-reduce): the
// block min-reduction of this partition, then thread r writes the candidate into partition r's mailbox slot
// (one 8-byte store over NVLink: round << 32 | time) and waits until partition r's candidate for the same round has
// arrived in the own mailbox. GVT = min over all. Every partition's appends of the previous round precede its slot
// write (kernel boundary + system fence), so once the slots are in, the mailbox records are complete: the exchange
// doubles as the barrier. One process per GPU: all partitions run this kernel at the same time, each on its own
// GPU; the wait is bounded and raises E_COMM instead of hanging.
// ------------------------------------------------------------------------------------------------
__global__ void k_gvt(View v) {
  __shared__ u64 sm[32];
  __shared__ u32 s_t[MAX_WORLD];
  Ctl* c = v.ctl;
  block_local_min(v, c, sm);
  __syncthreads();
  const u32 par = (c->round + 1u) & 1u;
  const u32 R = c->round + 1u + (c->epoch << 24);      // stamp of the round that is about to start
  const u64 lm = c->local_min;
  const u32 my_t = lm == KEY_MAX ? 0xFFFFFFFFu : key_time(lm);
  if (threadIdx.x < v.world) {
    const u32 r = threadIdx.x;
    u32 t = my_t;
    if (r != v.rank) {
      __threadfence_system();
      *reinterpret_cast<volatile u64*>(mb_gvt(v, r, par, v.rank)) = ((u64)R << 32) | my_t;
      const volatile u64* mine = mb_gvt(v, v.rank, par, r);
      u64 x = *mine;
      for (u32 spin = 0; (u32)(x >> 32) != R && spin < (1u << 27); ++spin) { __nanosleep(20); x = *mine; }
      if ((u32)(x >> 32) != R) { atomicOr(&c->error, E_COMM); t = 0xFFFFFFFFu; }
      else t = (u32)x;
      __threadfence_system();
    }
    s_t[r] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0xFFFFFFFFu;
    for (u32 r = 0; r < v.world; ++r) t = s_t[r] < t ? s_t[r] : t;
    c->gvt = t == 0xFFFFFFFFu ? KEY_MAX : make_key(t, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// k_plan — GVT := (all-reduced) local_min. Finish test of runner.hpp:392. Commit list = buckets entirely below GVT
// (R8/R9). Gather list = unconsumed tails of the buckets inside the optimistic window
// [bucket(GVT), bucket(GVT) + window). The control block is staged in shared memory (one coalesced load, one
// coalesced store): the serial bookkeeping of thread 0 then costs no global round trips. Nothing else touches the
// control block while k_plan runs (it is the first kernel of a round).
// ------------------------------------------------------------------------------------------------
__global__ void k_plan(View v, int use_global_gvt) {
  __shared__ u64 sm[32];
  __shared__ __align__(16) u32 s_words[(sizeof(Ctl) + 3) / 4];
  Ctl* c = reinterpret_cast<Ctl*>(s_words);
  constexpr u32 kWords = (u32)(sizeof(Ctl) / 4);
  static_assert(sizeof(Ctl) % 4 == 0, "Ctl is copied word-wise");
  for (u32 w = threadIdx.x; w < kWords; w += blockDim.x) s_words[w] = reinterpret_cast<const u32*>(v.ctl)[w];
  __syncthreads();
  if (!use_global_gvt) {     // single partition: the min-reduction is fused here
    block_local_min(v, c, sm);
  }
  if (threadIdx.x == 0) {
    u64 gvt = use_global_gvt ? c->gvt : c->local_min;
    c->gvt = gvt;
    u32 t = gvt == KEY_MAX ? 0xFFFFFFFFu : key_time(gvt);
    bool done = t >= v.finish_time;
    u32 fin_b = (v.finish_time + v.bucket_ticks - 1) / v.bucket_ticks;   // buckets [0, fin_b) hold times < finish
    if (fin_b > v.nb) fin_b = v.nb;
    u32 cur = done ? fin_b : t / v.bucket_ticks;
    if (cur > fin_b) cur = fin_b;
    c->cur_abs = cur;
    // commit list (history replay commits nothing: the loaded events stay in the calendar as processed events)
    u32 nc = 0, tot = 0;
    if (!v.replay) {
      for (u32 b = c->commit_next; b < cur; ++b) {
        u32 f = v.fill[b];
        if (f) { v.commit[nc++] = make_uint4(b, f, tot, 0); tot += f; }
      }
      c->commit_next = cur > c->commit_next ? cur : c->commit_next;
    }
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
        const u32 cons = v.consumed[b];
        u32 avail = v.fill[b] - cons;
        if (!avail) continue;
        u32 take = avail < v.cap_batch - total ? avail : v.cap_batch - total;
        if (take) {
          v.gather[ng++] = make_uint4(b, cons, take, total);
          total += take;
          v.consumed[b] = cons + take;
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
    c->n_vheavy = 0;
    c->n_fix = 0;
    c->done = done ? 1u : 0u;
    c->log_stable = c->log_count;      // the commit passes of all earlier rounds are complete
    c->round++;
    for (int i = 0; i < MAX_WORLD; ++i) c->send_count[i] = 0;
    c->sent_min = KEY_MAX;
    if (v.use_graph) {
      // GVT must keep advancing (every round commits or integrates something): give up instead of spinning forever
      if (gvt == c->prev_gvt) { if (++c->stall > kStallRounds) c->error |= E_STALL; } else { c->stall = 0; c->prev_gvt = gvt; }
      cudaGraphSetConditional(v.h_loop, (done || c->error) ? 0u : 1u);
    }
    if (v.p2p) {
      // the mailbox half that was just integrated (parity of the round that ended) is free again
      const u32 par = (c->round - 1) & 1u;
      for (u32 r = 0; r < v.world; ++r) { *mb_cnt(v, v.rank, par, r, 0) = 0; *mb_cnt(v, v.rank, par, r, 1) = 0; }
    }
  }
  __syncthreads();
  for (u32 w = threadIdx.x; w < kWords; w += blockDim.x) reinterpret_cast<u32*>(v.ctl)[w] = s_words[w];
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
    u32 my_slot = 0;
    if (i < total) {
      u32 lo = 0, hi = nc;        // locate the bucket: binary search over the commit list offsets
      while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (v.commit[mid].z <= i) lo = mid; else hi = mid;
      }
      uint4 ce = v.commit[lo];
      u32 slot = slot_of(v, ce.x, i - ce.z);
      my_slot = slot;
      e = load_ev_ef(v.ev + slot);
      const unsigned char pf = v.pflag[slot];
      // an event loaded from the history store is output again only if the repeat run re-executed it (R13)
      keep = !(e.flags & F_CANCEL) && !(pf & (P_DEAD | P_HISTORY)) && key_time(e.key) < v.finish_time;
      if (keep) {
        if (!(pf & P_PROCESSED)) internal_error(v, 1, key_time(e.key));
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
        if (pos < v.cap_log) {
          e.flags = tg;
          store_ev_ef(v.log + pos, e);
          if (v.log_sent) {      // --diff_init: the sent-log entry goes to the history store with the event (R12)
            const Meta m = load_meta(v.meta + my_slot);
            v.log_sent[pos] = make_uint4(m.sent_dst, (u32)m.sent_key, (u32)(m.sent_key >> 32), 0u);
          }
        }
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
// calendar slot of the arrival with index i (arrival order = calendar order of the gathered bucket tails)
// ------------------------------------------------------------------------------------------------
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
// k_link — arrivals of this window, in calendar (slot) order: one coalesced pass over the event records. Replaces
// the inbox append of lp::buffer (logical_process.hpp:115-127) and scheduler::queue (process_scheduler.hpp:55-81,
// "which LPs have work"): every arrival is counted on its LP and chained to the LP's previous arrival of the round
// (atomic exchange on the LP's list head), so that afterwards any arrival can see all arrivals of its LP without
// the events being moved or sorted. The first arrival of an LP notes the LP's processed-chain head as it was before
// the round; the arrival that makes the LP's count exceed kLightMax puts the LP on the long-segment list.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_link(View v) {
  Ctl* c = v.ctl;
  const u32 n = c->n_arr;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of_arrival(v, i);
    const Ev e = load_ev(v.ev + slot);
    const u32 lp = part_local(v, e.dst);
    u32* h = reinterpret_cast<u32*>(v.lph + lp);
    const u32 rank = atomicAdd(h, 1u) & kCntMask;
    const u32 prev1 = atomicExch(h + 1, i + 1u);
    const u32 prev = prev1 ? prev1 - 1u : NIL31;
    v.link[i] = make_uint4((u32)e.key, (u32)(e.key >> 32), prev | ((e.flags & F_CANCEL) ? 0x80000000u : 0u), slot);
    v.tmp_rank[i] = rank;
    if (rank == 0) *reinterpret_cast<uint2*>(h + 2) = v.last[lp];
    if (rank == v.light_max) v.heavy[atomicAdd(&c->n_heavy, 1u)] = make_uint4(lp, 0, 0, 0);
  }
}

// k_segs_heavy — the LPs with more than kLightMax arrivals get a segment of `sorted` (one atomic each).
__global__ void k_segs_heavy(View v) {
  Ctl* c = v.ctl;
  const u32 nh = c->n_heavy;
  for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x) {
    const u32 lp = v.heavy[h].x;
    const u32 n = v.lph[lp].x & kCntMask;
    const u32 start = atomicAdd(&c->seg_cursor, n);
    v.lph[lp].y = start;          // the list head is not needed for a long LP: the word now holds its segment start
    v.heavy[h] = make_uint4(lp, start, n, 0);
    if (n > kRegSortSmall) atomicAdd(&c->n_vheavy, 1u);
  }
}

// first flagger of an LP puts it on the fix list
__device__ __forceinline__ void flag_lp(const View& v, u32 lp, u32 bit) {
  if (v.replay) internal_error(v, 2, lp);      // a recorded history has no irregular turns
  const u32 old = atomicOr(reinterpret_cast<u32*>(v.lph + lp), bit << kCntBits);
  if (((old >> kCntBits) & PL_FIX) == 0) v.fix[atomicAdd(&v.ctl->n_fix, 1u)] = lp;
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
__device__ __forceinline__ Sorted load_sorted(const Sorted* p) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  Sorted x;
  x.key = ((u64)q.y << 32) | q.x; x.slot = q.z; x.seq = q.w;
  return x;
}
__device__ __forceinline__ void store_sorted(Sorted* p, const Sorted& x) {
  *reinterpret_cast<uint4*>(p) = make_uint4((u32)x.key, (u32)(x.key >> 32), x.slot, x.seq);
}

template <class M>
struct Turn {
  const View& v;
  const typename M::Data& md;
  u32 last_slot, last_time;
  u32 st[M::kStateWords];
  bool st_dirty;
  u32 err;
  u32 n_proc, n_undone, n_anti, n_annih, n_dup, n_requeued;
  u32 tbound;      // first time beyond the current window

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

  // An undone event beyond the current window goes back to PENDING instead of being re-executed at once: a copy is
  // appended to its calendar bucket (through the outbox) and the recorded slot dies. What lp::flush_buf does by
  // lowering local_time_ (logical_process.hpp:145-156): the rolled-back events wait in the queue until the
  // simulation reaches them again. Re-executing them eagerly would let every later arrival at this LP roll the
  // same future back again (quadratic in a --diff_repeat cascade).
  __device__ __forceinline__ void requeue(u32 slot) {
    Ev e = load_ev(v.ev + slot);
    e.flags = 0;
    store_ev(v.outbox + extra_idx(), e);
    v.pflag[slot] = P_DEAD;
    ++n_requeued;
  }

  __device__ __forceinline__ u32 extra_idx() {
    u32 pos = atomicAdd(&v.ctl->n_extra, 1u);
    if (pos >= v.cap_aux) { err |= E_BATCH; pos = v.cap_aux - 1; }
    return v.cap_batch + pos;
  }

  // seg[0..n) = the LP's arrivals of this round, sorted; old_slot/old_time = the LP's chain head before the round.
  // Arrival i owns outbox entry i, so whatever k_exec may already have written for this LP is overwritten here.
  __device__ void run(u32 lp, const Sorted* seg, u32 n, u32 old_slot, u32 old_time) {
    err = 0; n_proc = n_undone = n_anti = n_annih = n_dup = n_requeued = 0;
    tbound = v.ctl->bound_abs * v.bucket_ticks;
    st_dirty = false;
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
        const Sorted a = load_sorted(seg + ia);
        const u32 my_out = a.seq & 0x7FFFFFFFu;
        const u32 a_slot = a.slot;
        if (a.seq & 0x80000000u) {
          v.pflag[a_slot] = P_DEAD;                     // the anti-message itself is consumed
          v.outbox[my_out].key = KEY_MAX;
          if (present != NIL) {                         // erase (queue.hpp:87-90)
            v.pflag[present] = P_DEAD;
            present = NIL;
            if (out_idx != NIL) { v.outbox[out_idx].key = KEY_MAX; out_idx = NIL; }
            ++n_annih;
          }
        } else if (present == NIL) {
          present = a_slot;                             // insert (queue.hpp:92)
          out_idx = my_out;
        } else {
          v.pflag[a_slot] = P_DEAD;                     // duplicate key: second insert ignored
          v.outbox[my_out].key = KEY_MAX;
          ++n_dup;
        }
        ++ia;
      }
      if (present != NIL) {
        // a re-executed (rolled back) event has no inbox position: its output goes to the extra region
        if (key_time(k) < tbound) process(present, out_idx != NIL ? out_idx : extra_idx());
        else requeue(present);
      }
    }

    v.last[lp] = make_uint2(last_slot, last_time);
    if (st_dirty) {
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) v.state[(size_t)lp * M::kStateWords + w] = st[w];
    }
  }
};

// ------------------------------------------------------------------------------------------------
// exec_one — the plain case of an LP turn for ONE event: run the handler on the LP's current state, link the event
// into the LP's processed chain behind its predecessor in key order, record the sent-log entry, write the produced
// event to outbox entry i. This is the speculative form of runner::run_lp_aggr (runner.hpp:530-570): all events of
// an LP are executed against the same state, which is exactly the sequential result as long as no invocation changes
// the state (neither shipped application ever does). The moment one does, the LP is flagged and k_fixup re-executes
// its whole turn sequentially — handlers that mutate state keep exact semantics.
// ------------------------------------------------------------------------------------------------
template <class M, bool kLinkLater = false>
__device__ __forceinline__ void exec_one(const View& v, const typename M::Data& md, u32 i, u32 slot, Ev e, u32 lp,
                                         u32 pred_slot, u32 pred_time, bool is_last, u32& err) {
  e.flags = 0;
  if (v.replay) {
    // history replay (--diff_repeat, R13): the event was processed in the recorded run; its sent-log entry came with
    // it. Only its place in the LP's processed chain is established, nothing is executed or sent.
    Meta m = load_meta(v.meta + slot);
    m.link = pred_slot;
    m.prevtime = pred_time;
    store_meta(v.meta + slot, m);
    v.pflag[slot] = P_PROCESSED | P_HISTORY;
    v.outbox[i].key = KEY_MAX;
    if (is_last) v.last[lp] = make_uint2(slot, key_time(e.key));
    return;
  }
  u32 st[M::kStateWords], pre[M::kStateWords];
#pragma unroll
  for (int q = 0; q < M::kStateWords; ++q) { st[q] = v.state[(size_t)lp * M::kStateWords + q]; pre[q] = st[q]; }
  Ev out;
  const bool has = M::handle(md, e, st, out, err);
  bool changed = false;
#pragma unroll
  for (int q = 0; q < M::kStateWords; ++q) changed |= (pre[q] != st[q]);
  if (has && changed) { flag_lp(v, lp, PL_MUTATED); return; }   // state-mutating handler
  Meta m;
  m.link = pred_slot;
  m.prevtime = pred_time;
  m.sent_dst = NIL;
  m.snap = NIL;
  m.sent_key = 0;
  if (has) {
    m.sent_dst = out.dst;
    m.sent_key = out.key;
    store_ev(v.outbox + i, out);
  } else {
    v.outbox[i].key = KEY_MAX;
  }
  // a meta record is not read again unless its LP rolls back: evict-first. Exception: the record of a long LP's event,
  // whose chain link k_exec_heavy fills in right afterwards.
  if (kLinkLater) store_meta(v.meta + slot, m); else store_meta_ef(v.meta + slot, m);
  v.pflag[slot] = P_PROCESSED;
  if (is_last) v.last[lp] = make_uint2(slot, key_time(e.key));
}

// ------------------------------------------------------------------------------------------------
// k_exec — one thread per EVENT of the window, in calendar (slot) order, so the event read and the meta / pflag /
// outbox writes are coalesced streams; only the per-LP words are gathered. Skew-free: a hot LP's events are spread
// over many threads.
//   * LP with up to kLightMax arrivals: the thread walks the LP's arrival list (link records) to find its place in
//     the LP's key order (R1): predecessor = largest smaller key, "last" = no larger key. The walk also finds the
//     irregular turns — an anti-message, a duplicate key, or a key not later than the LP's newest processed event
//     (straggler, R5) — which are left to the sequential lp_turn of k_fixup.
//   * LP with more arrivals: the thread only drops (key, slot, arrival) into the LP's segment of `sorted`;
//     k_sortheavy sorts the segment and k_exec_heavy executes it.
// ------------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(256) k_exec(View v, typename M::Data md) {
  Ctl* c = v.ctl;
  const u32 n_arr = c->n_arr;
  u32 err = 0;
  if (v.use_graph && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(v.h_vheavy, c->n_vheavy ? 1u : 0u);
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_arr; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of_arrival(v, i);
    const Ev e = load_ev(v.ev + slot);
    const u32 lp = part_local(v, e.dst);
    const uint4 h = v.lph[lp];
    const u32 n = h.x & kCntMask;
    const u32 anti = (e.flags & F_CANCEL) ? 0x80000000u : 0u;
    if (n > v.light_max) {
      // long LP: drop (key, slot, arrival) into the LP's segment and execute speculatively right here, where the
      // event, meta and outbox accesses are coalesced; k_exec_heavy adds the processed-chain link afterwards
      Sorted x;
      x.key = e.key; x.slot = slot; x.seq = i | anti;
      store_sorted(v.sorted + h.y + v.tmp_rank[i], x);
      if (anti) { flag_lp(v, lp, PL_IRREGULAR); continue; }
      if ((h.x >> kCntBits) & PL_FIX) continue;
      exec_one<M, true>(v, md, i, slot, e, lp, NIL, 0u, false, err);
      continue;
    }
    if ((h.x >> kCntBits) & PL_FIX) continue;             // handled by k_fixup
    u32 pred_slot = h.z, pred_time = h.w;
    bool is_last = true;
    bool bad = anti != 0 || (h.z != NIL && key_time(e.key) <= h.w);
    if (n > 1) {
      u64 pk = 0;
      bool have = false;
      u32 k = h.y - 1u;
      for (u32 step = 0; step < n && k != NIL31; ++step) {
        const uint4 L = v.link[k];
        if (k != i) {
          const u64 tk = ((u64)L.y << 32) | L.x;
          if (L.z >> 31) bad = true;
          if (tk == e.key) bad = true;
          else if (tk < e.key) {
            if (!have || tk > pk) { pk = tk; pred_slot = L.w; have = true; }
          } else {
            is_last = false;
          }
        }
        k = L.z & 0x7FFFFFFFu;
      }
      if (have) pred_time = key_time(pk);
    }
    if (bad) { flag_lp(v, lp, PL_IRREGULAR); continue; }
    exec_one<M>(v, md, i, slot, e, lp, pred_slot, pred_time, is_last, err);
  }
  if (err) atomicOr(&c->error, err);
}

// ------------------------------------------------------------------------------------------------
// k_sortheavy — segments longer than kRegSortSmall (the very hot LPs: at N GPUs PHOLD's hot LPs hold N times as
// many events): one BLOCK of 256 threads per LP sorts the segment by (key, arrival order) (R1) and classifies it: an
// anti-message, a duplicate key or a possible straggler makes the turn irregular (-> k_fixup); a plain turn gets its
// processed-chain links right here, as in k_exec_heavy.
// Hybrid bitonic network, R = 1, 2, 4 or 8 entries per thread in registers (entry index = tid + 256 r):
//   partner distance < 32        register shuffle inside the warp
//   32 <= distance < 256         exchange through a 4 KB shared-memory tile (the only shared-memory traffic:
//                                3 of the log2(P) steps of a merge level — a pure shared-memory network is bound by
//                                shared-memory bandwidth, measured 257 us per window for 10 500 segments of ~200)
//   distance >= 256              the other register of the same thread
// Longer than kBlockSortCap: one thread sorts in place (not reached by any shipped workload).
// ------------------------------------------------------------------------------------------------
constexpr u32 kBlockSortCap = 2048;
constexpr int kSortThreads = 256;

struct KS { u64 key; u32 slot, seq; };      // sort element: key, calendar slot, arrival index | anti << 31
__device__ __forceinline__ bool ks_less(const KS& a, const KS& b) {
  return a.key < b.key || (a.key == b.key && (a.seq & 0x7FFFFFFFu) < (b.seq & 0x7FFFFFFFu));
}
__device__ __forceinline__ KS ks_shfl_xor(const KS& x, int m) {
  KS y;
  const u32 lo = __shfl_xor_sync(0xffffffffu, (u32)x.key, m), hi = __shfl_xor_sync(0xffffffffu, (u32)(x.key >> 32), m);
  y.key = ((u64)hi << 32) | lo;
  y.slot = __shfl_xor_sync(0xffffffffu, x.slot, m);
  y.seq = __shfl_xor_sync(0xffffffffu, x.seq, m);
  return y;
}

template <int R>
__device__ __forceinline__ void block_sort_segment(const View& v, u32 lp, Sorted* seg, u32 n, uint4* tile, u64* s_last,
                                                   u32* s_lslot, u32* s_flag, u64& s_first) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KS x[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u32 e = tid + ((u32)r << 8);
    if (e < n) { const uint4 q = *reinterpret_cast<const uint4*>(seg + e); x[r].key = ((u64)q.y << 32) | q.x; x[r].slot = q.z; x[r].seq = q.w; }
    else { x[r].key = KEY_MAX; x[r].slot = NIL; x[r].seq = 0x7FFFFFFFu; }
  }
  if (tid == 0) *s_flag = 0;
#pragma unroll
  for (u32 k2 = 2; k2 <= 256u * R; k2 <<= 1) {
#pragma unroll
    for (u32 j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      if (j2 >= 256u) {
        const u32 dr = j2 >> 8;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & dr) == 0) {
            const bool asc = (((u32)r << 8) & k2) == 0;
            const bool sw = asc ? ks_less(x[r | dr], x[r]) : ks_less(x[r], x[r | dr]);
            if (sw) { const KS t = x[r]; x[r] = x[r | dr]; x[r | dr] = t; }
          }
        }
      } else if (j2 >= 32u) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          tile[tid] = make_uint4((u32)x[r].key, (u32)(x[r].key >> 32), x[r].slot, x[r].seq);
          __syncthreads();
          const uint4 q = tile[tid ^ j2];
          __syncthreads();
          KS y;
          y.key = ((u64)q.y << 32) | q.x; y.slot = q.z; y.seq = q.w;
          const u32 e = tid + ((u32)r << 8);
          const bool keep_min = ((tid & j2) == 0) == ((e & k2) == 0);
          if (keep_min ? ks_less(y, x[r]) : ks_less(x[r], y)) x[r] = y;
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const KS y = ks_shfl_xor(x[r], (int)j2);
          const u32 e = tid + ((u32)r << 8);
          const bool keep_min = ((lane & j2) == 0) == ((e & k2) == 0);
          if (keep_min ? ks_less(y, x[r]) : ks_less(x[r], y)) x[r] = y;
        }
      }
    }
  }
  // every entry looks at the entry before it: duplicate key, and its predecessor in the LP's processed chain
#pragma unroll
  for (int r = 0; r < R; ++r) if (lane == 31) { s_last[r * 8 + warp] = x[r].key; s_lslot[r * 8 + warp] = x[r].slot; }
  __syncthreads();
  const uint4 h = v.lph[lp];
  const bool fixed = ((h.x >> kCntBits) & PL_FIX) != 0;      // k_exec already sent the LP to k_fixup: only sort
  u32 flag = 0;
  u64 pkey[R];
  u32 pslot[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u32 e = tid + ((u32)r << 8);
    const u32 lo = __shfl_up_sync(0xffffffffu, (u32)x[r].key, 1), hi = __shfl_up_sync(0xffffffffu, (u32)(x[r].key >> 32), 1);
    pkey[r] = ((u64)hi << 32) | lo;
    pslot[r] = __shfl_up_sync(0xffffffffu, x[r].slot, 1);
    if (lane == 0 && e) {
      const u32 w = warp ? r * 8 + warp - 1 : (r - 1) * 8 + 7;
      pkey[r] = s_last[w]; pslot[r] = s_lslot[w];
    }
    if (e < n) {
      flag |= x[r].seq >> 31;
      if (e && pkey[r] == x[r].key) flag = 1;
    }
  }
  if (tid == 0) s_first = x[0].key;
  if (flag) *s_flag = 1;
  __syncthreads();
  bool irregular = *s_flag != 0;
  if (h.z != NIL && key_time(s_first) <= h.w) irregular = true;      // possible straggler
  if (irregular || fixed) {
    // leave the sorted segment for the sequential lp_turn of k_fixup
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const u32 e = tid + ((u32)r << 8);
      if (e < n) *reinterpret_cast<uint4*>(seg + e) = make_uint4((u32)x[r].key, (u32)(x[r].key >> 32), x[r].slot, x[r].seq);
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicOr(reinterpret_cast<u32*>(v.lph + lp), PL_SORTED << kCntBits);
      if (!fixed) flag_lp(v, lp, PL_IRREGULAR);
    }
  } else {
    // plain turn: k_exec has run the handlers; add every event's place in the processed chain
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const u32 e = tid + ((u32)r << 8);
      if (e < n) {
        u32 ps = pslot[r], ptime = key_time(pkey[r]);
        if (e == 0) { ps = h.z; ptime = h.w; }
        *reinterpret_cast<uint2*>(v.meta + x[r].slot) = make_uint2(ps, ptime);
        if (e + 1u == n) v.last[lp] = make_uint2(x[r].slot, key_time(x[r].key));
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSortThreads) k_sortheavy(View v) {
  __shared__ uint4 tile[kSortThreads];
  __shared__ u64 s_last[8 * 8];
  __shared__ u32 s_lslot[8 * 8];
  __shared__ u32 s_flag;
  __shared__ u64 s_first;
  Ctl* c = v.ctl;
  const u32 nh = c->n_vheavy ? c->n_heavy : 0u;      // n_vheavy counts the segments longer than kRegSortSmall
  if (nh == 0) return;
  for (u32 h = blockIdx.x; h < nh; h += gridDim.x) {
    const uint4 a = v.heavy[h];
    const u32 lp = a.x, n = a.z;
    if (n <= kRegSortSmall) continue;
    Sorted* seg = v.sorted + a.y;
    if (n <= 256) block_sort_segment<1>(v, lp, seg, n, tile, s_last, s_lslot, &s_flag, s_first);
    else if (n <= 512) block_sort_segment<2>(v, lp, seg, n, tile, s_last, s_lslot, &s_flag, s_first);
    else if (n <= 1024) block_sort_segment<4>(v, lp, seg, n, tile, s_last, s_lslot, &s_flag, s_first);
    else if (n <= kBlockSortCap) block_sort_segment<8>(v, lp, seg, n, tile, s_last, s_lslot, &s_flag, s_first);
    else {
      if (threadIdx.x == 0) {
        u32 flag = 0;
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
        for (u32 i = 0; i < n; ++i) {
          flag |= seg[i].seq >> 31;
          if (i && seg[i - 1].key == seg[i].key) flag = 1;
        }
        const uint4 hh = v.lph[lp];
        bool irregular = flag != 0 || (hh.z != NIL && key_time(seg[0].key) <= hh.w);
        __threadfence();
        atomicOr(reinterpret_cast<u32*>(v.lph + lp), PL_SORTED << kCntBits);
        if ((hh.x >> kCntBits) & PL_FIX) irregular = false;      // already on the fix list
        else if (irregular) flag_lp(v, lp, PL_IRREGULAR);
        if (!irregular && !((hh.x >> kCntBits) & PL_FIX)) {
          u32 ps = hh.z, pt = hh.w;
          for (u32 i = 0; i < n; ++i) {
            *reinterpret_cast<uint2*>(v.meta + seg[i].slot) = make_uint2(ps, pt);
            ps = seg[i].slot; pt = key_time(seg[i].key);
          }
          v.last[lp] = make_uint2(ps, pt);
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_exec_heavy — one WARP per long LP (the hot LPs). k_exec has already run the handler for the LP's events, in the
// coalesced pass; what is left is each event's place in the LP's key order (R1). The warp loads the segment
// (key, slot, arrival) into registers — R entries per lane, R = 1, 2 or 4 — sorts it with a bitonic network over register
// shuffles, and each entry reads its predecessor off its neighbour: processed-chain link -> 8-byte store into the
// event's meta record, the last entry becomes the LP's chain head. An anti-message, a duplicate key or a possible
// straggler makes the turn irregular: the sorted segment is written back for the sequential lp_turn of k_fixup
// (also when k_exec flagged the LP because its handler changed the state). Segments longer than kRegSortSmall are
// k_sortheavy's (one block per LP).
// ------------------------------------------------------------------------------------------------
// entry index = lane + 32 * r. Sorts 32 * R entries ascending.
template <int R>
__device__ __forceinline__ void warp_bitonic(KS (&x)[R], u32 lane) {
#pragma unroll
  for (u32 k2 = 2; k2 <= 32u * R; k2 <<= 1) {
#pragma unroll
    for (u32 j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      if (j2 >= 32u) {
        const u32 dr = j2 >> 5;                 // partner register
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & dr) == 0) {
            const bool asc = (((u32)r << 5) & k2) == 0;       // lane bits are below bit 5, k2 >= 64 here
            const bool sw = asc ? ks_less(x[r | dr], x[r]) : ks_less(x[r], x[r | dr]);
            if (sw) { const KS t = x[r]; x[r] = x[r | dr]; x[r | dr] = t; }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const KS y = ks_shfl_xor(x[r], (int)j2);
          const u32 e = lane + ((u32)r << 5);
          const bool keep_min = ((lane & j2) == 0) == ((e & k2) == 0);
          if (keep_min ? ks_less(y, x[r]) : ks_less(x[r], y)) x[r] = y;
        }
      }
    }
  }
}

template <int R>
__device__ __forceinline__ void heavy_turn(const View& v, u32 lp, u32 start, u32 n, const uint4 h, u32 lane) {
  const bool fixed = ((h.x >> kCntBits) & PL_FIX) != 0;      // k_exec already sent the LP to k_fixup: only sort
  KS x[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u32 i = lane + ((u32)r << 5);
    if (i < n) { const uint4 q = *reinterpret_cast<const uint4*>(v.sorted + start + i); x[r].key = ((u64)q.y << 32) | q.x; x[r].slot = q.z; x[r].seq = q.w; }
    else { x[r].key = KEY_MAX; x[r].slot = NIL; x[r].seq = 0x7FFFFFFFu; }
  }
  warp_bitonic<R>(x, lane);
  // neighbours: the entry before (index - 1) of every entry
  u32 flag = 0;
  u64 pkey[R];
  u32 pslot[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    u32 lo = __shfl_up_sync(0xffffffffu, (u32)x[r].key, 1), hi = __shfl_up_sync(0xffffffffu, (u32)(x[r].key >> 32), 1);
    u32 sl = __shfl_up_sync(0xffffffffu, x[r].slot, 1);
    if (r > 0) {      // lane 0 takes the last lane's entry of the previous register
      const u32 lo2 = __shfl_sync(0xffffffffu, (u32)x[r - 1].key, 31), hi2 = __shfl_sync(0xffffffffu, (u32)(x[r - 1].key >> 32), 31);
      const u32 sl2 = __shfl_sync(0xffffffffu, x[r - 1].slot, 31);
      if (lane == 0) { lo = lo2; hi = hi2; sl = sl2; }
    }
    pkey[r] = ((u64)hi << 32) | lo;
    pslot[r] = sl;
    const u32 i = lane + ((u32)r << 5);
    if (i < n) {
      flag |= x[r].seq >> 31;
      if (i && pkey[r] == x[r].key) flag = 1;
    }
  }
  const u64 first_key = ((u64)__shfl_sync(0xffffffffu, (u32)(x[0].key >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (u32)x[0].key, 0);
  bool irregular = __any_sync(0xffffffffu, flag != 0);
  if (h.z != NIL && key_time(first_key) <= h.w) irregular = true;      // possible straggler
  if (irregular || fixed) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const u32 i = lane + ((u32)r << 5);
      if (i < n) *reinterpret_cast<uint4*>(v.sorted + start + i) = make_uint4((u32)x[r].key, (u32)(x[r].key >> 32), x[r].slot, x[r].seq);
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicOr(reinterpret_cast<u32*>(v.lph + lp), PL_SORTED << kCntBits);
      if (!fixed) flag_lp(v, lp, PL_IRREGULAR);
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u32 i = lane + ((u32)r << 5);
    if (i < n) {
      u32 ps = pslot[r], ptime = key_time(pkey[r]);
      if (i == 0) { ps = h.z; ptime = h.w; }
      *reinterpret_cast<uint2*>(v.meta + x[r].slot) = make_uint2(ps, ptime);
      if (i + 1u == n) v.last[lp] = make_uint2(x[r].slot, key_time(x[r].key));
    }
  }
}

__global__ void __launch_bounds__(256) k_exec_heavy(View v) {
  Ctl* c = v.ctl;
  const u32 nh = c->n_heavy;
  const u32 lane = threadIdx.x & 31;
  const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 hi = wid; hi < nh; hi += nw) {
    const uint4 a = v.heavy[hi];
    const u32 lp = a.x, start = a.y, n = a.z;
    if (n > kRegSortSmall) continue;                  // k_sortheavy's
    const uint4 h = v.lph[lp];
    if (n <= 32) heavy_turn<1>(v, lp, start, n, h, lane);
    else if (n <= 64) heavy_turn<2>(v, lp, start, n, h, lane);
    else heavy_turn<4>(v, lp, start, n, h, lane);
  }
  if (v.use_graph) {
    // the last block to finish knows the final fix list: it opens or closes the IF node around k_fixup
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&c->heavy_blocks, 1u) == gridDim.x - 1) {
        c->heavy_blocks = 0;
        cudaGraphSetConditional(v.h_fix, *reinterpret_cast<volatile u32*>(&c->n_fix) ? 1u : 0u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fixup — the LPs whose turn is not plain (anti-message, duplicate, straggler, state-mutating handler): one thread
// per LP runs the exact sequential lp_turn. Long segments were sorted by k_sortheavy; a short one is collected
// from the LP's arrival list and sorted here.
// ------------------------------------------------------------------------------------------------
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
      const u32 lp = v.fix[k];
      const uint4 h = v.lph[lp];
      const u32 n = h.x & kCntMask;
      u32 s0;
      if ((h.x >> kCntBits) & PL_SORTED) {
        s0 = h.y;
      } else {
        Sorted loc[kLightMax];
        u32 m = 0;
        for (u32 a = h.y - 1u; a != NIL31 && m < kLightMax; ++m) {
          const uint4 L = v.link[a];
          loc[m].key = ((u64)L.y << 32) | L.x; loc[m].slot = L.w; loc[m].seq = a | (L.z & 0x80000000u);
          a = L.z & 0x7FFFFFFFu;
        }
        if (m != n) internal_error(v, 3, lp);
        for (u32 q = 1; q < m; ++q) {
          Sorted x = loc[q];
          u32 j = q;
          while (j > 0 && sorted_less(x, loc[j - 1])) { loc[j] = loc[j - 1]; --j; }
          loc[j] = x;
        }
        s0 = atomicAdd(&c->seg_cursor, m);
        for (u32 q = 0; q < m; ++q) store_sorted(v.sorted + s0 + q, loc[q]);
      }
      n_redo += n;
      Turn<M> t(v, md);
      t.run(lp, v.sorted + s0, n, h.z, h.w);
      n_proc += t.n_proc; n_undone += t.n_undone; n_anti += t.n_anti; n_annih += t.n_annih; n_dup += t.n_dup;
      err |= t.err;
    }
  }
  __syncwarp();
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
  if (ch == NIL) internal_error(v, 4, 0);
  return ch;
}

// --- TMA 1-D bulk copy global -> shared, completion on an mbarrier (sm_90+/sm_100a) ---
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Each block stages one tile of the source (a contiguous run of 32-byte records) in shared memory with ONE TMA bulk
// copy, builds the per-target histogram from the staged records, claims one range per target with a single global
// atomic, and scatters the records from shared memory with 256-bit stores. Tiles are sized so that every block of
// the persistent grid gets the same number of equally large tiles (no tail wave).
__global__ void __launch_bounds__(kRouteThreads) k_route(View v, int which, u32 peer, u32 filter) {
  extern __shared__ __align__(128) unsigned char smem_route[];
  __shared__ u64 s_bar;
  __shared__ u64 s_sent_min;
  __shared__ const Ev* s_seg[MAX_WORLD + 2];     // the source = a concatenation of up to world + 1 record runs
  __shared__ u32 s_off[MAX_WORLD + 3];
  __shared__ u32 s_nseg;
  Ctl* c = v.ctl;
  if (threadIdx.x == 0) {
    u32 ns = 0, tot = 0;
    auto add = [&](const Ev* p, u32 n) { if (n) { s_seg[ns] = p; s_off[ns] = tot; tot += n; ++ns; } };
    if (which == 0) add(v.anti_box, c->n_anti < v.cap_aux ? c->n_anti : v.cap_aux);
    else if (which == 1) { add(v.outbox, c->n_arr); add(v.outbox + v.cap_batch, c->n_extra < v.cap_aux ? c->n_extra : v.cap_aux); }
    else if (which == 2) {
      if (v.p2p) {
        // own mailbox, all sources; `peer` = kind (1 = anti-message regions, integrated first; 0 = positives)
        const u32 par = c->round & 1u, kind = peer, cap = kind ? v.mb_cap_anti[v.rank] : v.mb_cap_send[v.rank];
        for (u32 r = 0; r < v.world; ++r) {
          if (r == v.rank) continue;
          u32 n = *mb_cnt(v, v.rank, par, r, kind);
          if (n > cap) { n = cap; if (blockIdx.x == 0) atomicOr(&c->error, E_SENDBUF); }
          add(mb_data(v, v.rank, par, r, kind), n);
        }
      } else {
        add(v.recvbuf + (size_t)peer * v.cap_send, c->recv_count[peer]);
      }
    }
    else add(v.outbox, c->load_count);
    s_off[ns] = tot;
    s_nseg = ns;
    s_sent_min = KEY_MAX;
    mbar_init(&s_bar, 1);
  }
  __syncthreads();
  const u32 total = s_off[s_nseg];
  if (total == 0) return;
  // equal tiles for all blocks: `passes` tiles of `tile` records each (tile is a multiple of 32 records = 1 KB)
  const u32 passes = (total + gridDim.x * kRouteTile - 1) / (gridDim.x * kRouteTile);
  u32 tile = (total + gridDim.x * passes - 1) / (gridDim.x * passes);
  tile = (tile + 31u) & ~31u;
  // shared-memory targets: [0, rs) = buckets base .. base+rs-1, rs = "beyond", rs+1+r = partition r.
  // Buckets further ahead than the table (tiny bucket_ticks) take one global atomic per event instead.
  const u32 rs = v.route_slots;
  const u32 nt = rs + 1 + v.world;
  const u32 base = c->commit_next;
  Ev* s_tile = reinterpret_cast<Ev*>(smem_route);                                  // [kRouteTile] staged records
  u32* s_cnt = reinterpret_cast<u32*>(smem_route + (size_t)kRouteTile * sizeof(Ev)); // [nt] events of the tile per target
  u32* s_base = s_cnt + nt;                             // [nt] first position claimed in the target
  u32* s_ch0 = s_base + nt;                             // [nt] chunk of that first position
  u32* s_ch1 = s_ch0 + nt;                              // [nt] the next chunk, if the range crosses into it
  const u32 chm = (1u << v.ch_log) - 1u;
  const u32 T_NONE = NIL, T_DIRECT = NIL - 1u;
  const u32 send_kind = which == 0 ? 1u : 0u;      // anti-messages travel in their own mailbox region
  u32 phase = 0;
  for (u32 t0 = blockIdx.x * tile; t0 < total; t0 += gridDim.x * tile) {
    const u32 cnt = total - t0 < tile ? total - t0 : tile;
    if (threadIdx.x == 0) {
      // one bulk copy per source run the tile overlaps
      mbar_expect_tx(&s_bar, cnt * (u32)sizeof(Ev));
      u32 sg = 0;
      while (s_off[sg + 1] <= t0) ++sg;
      for (u32 pos = t0, left = cnt, dst = 0; left; ++sg) {
        const u32 avail = s_off[sg + 1] - pos, take = avail < left ? avail : left;
        bulk_g2s(s_tile + dst, s_seg[sg] + (pos - s_off[sg]), take * (u32)sizeof(Ev), &s_bar);
        dst += take; pos += take; left -= take;
      }
    }
    for (u32 t = threadIdx.x; t < nt; t += kRouteThreads) s_cnt[t] = 0;
    __syncthreads();
    mbar_wait(&s_bar, phase);
    phase ^= 1u;
    u32 tgt[kRouteItems], rnk[kRouteItems];
    // pass 1: target of every event, rank inside the tile's share of the target
#pragma unroll
    for (int it = 0; it < kRouteItems; ++it) {
      const u32 i = it * kRouteThreads + threadIdx.x;
      tgt[it] = T_NONE;
      rnk[it] = 0;
      if (i < cnt) {
        const uint4 h = reinterpret_cast<const uint4*>(s_tile + i)[0];     // key lo, key hi, dst, src
        const u64 key = ((u64)h.y << 32) | h.x;
        bool take = key != KEY_MAX;
        if (take && filter) {
          const u32 fl = s_tile[i].flags;
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
            if (b < base) { internal_error(v, 5, (u32)which << 28 | b); continue; }   // causality: nothing may arrive below GVT
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
        if (v.p2p) s_base[t] = atomicAdd_system(mb_cnt(v, r, c->round & 1u, v.rank, send_kind), n);
        else s_base[t] = atomicAdd(&c->send_count[r], n);
        atomicAdd(&c->n_remote, (u64)n);
      }
    }
    // far buckets (beyond the table): the owner of a chunk's first position allocates it
#pragma unroll
    for (int it = 0; it < kRouteItems; ++it) {
      if (tgt[it] == T_DIRECT && (rnk[it] & chm) == 0) {
        const u32 i = it * kRouteThreads + threadIdx.x;
        const u32 b = key_time(s_tile[i].key) / v.bucket_ticks;
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
    // scatter the staged records to their slots
#pragma unroll
    for (int it = 0; it < kRouteItems; ++it) {
      const u32 t = tgt[it];
      if (t == T_NONE || t == rs) continue;
      const u32 i = it * kRouteThreads + threadIdx.x;
      const uint4* q = reinterpret_cast<const uint4*>(s_tile + i);
      const uint4 qa = q[0], qb = q[1];
      const u32 fl = qb.w & F_CANCEL;
      const u32 tm = qa.y;                       // key hi = receive_time
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
          ch = wait_chunk(v, &v.chunk_tab[(size_t)(tm / v.bucket_ticks) * v.maxc + (p >> v.ch_log)]);
        }
        if (ch == NIL) continue;
        const size_t slot = ((size_t)ch << v.ch_log) | (p & chm);
        st256_ef(v.ev + slot, qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, fl);
        v.pflag[slot] = 0;
        if (which == 3 && v.replay) {      // history load: the recorded sent-log entry becomes the event's meta record
          const uint4 sr = v.hist_sent[t0 + i];
          st256(v.meta + slot, NIL, 0u, sr.x, NIL, sr.y, sr.z, 0u, 0u);
        }
      } else {
        const u32 r = t - rs - 1;
        const u32 p = s_base[t] + rnk[it];
        if (p >= (v.p2p ? (send_kind ? v.mb_cap_anti[r] : v.mb_cap_send[r]) : v.cap_send)) { atomicOr(&c->error, E_SENDBUF); continue; }
        if (v.p2p) {
          // straight into the destination GPU's mailbox over NVLink
          st256(mb_data(v, r, c->round & 1u, v.rank, send_kind) + p, qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, fl);
          const u64 k = make_key(tm >= v.finish_time ? v.finish_time : tm / v.bucket_ticks * v.bucket_ticks, 0);
          if (k < s_sent_min) atomicMin(&s_sent_min, k);
        } else {
          st256(v.sendbuf + (size_t)r * v.cap_send + p, qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, fl);
        }
      }
    }
    __syncthreads();      // everyone is done with the staged tile before the next bulk copy overwrites it
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

// events loaded in the compact layout (ssb_load_events_raw): validated in place in the outbox staging area
__global__ void k_check_raw(View v, u32 n) {
  Ctl* c = v.ctl;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Ev e = load_ev(v.outbox + i);
    if (e.dst >= v.n_lp || key_time(e.key) >= 0x80000000u || key_id(e.key) == NIL) {
      atomicOr(&c->error, E_RANGE);
      v.outbox[i].key = KEY_MAX;
    } else if (e.flags & ~F_CANCEL) {
      v.outbox[i].flags = e.flags & F_CANCEL;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) c->load_count = n;
}

// committed log record -> ssb_packed_record (6 x u32)
__global__ void k_pack_log(const Ev* log, u64 first, u32 n, u32* out) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Ev e = load_ev(log + first + i);
    uint2* o = reinterpret_cast<uint2*>(out + (size_t)i * 6);
    o[0] = make_uint2(key_id(e.key), e.src);
    o[1] = make_uint2(e.send, e.dst);
    o[2] = make_uint2(key_time(e.key), e.flags);
  }
}

// history records (committed-log layout: flags = tag) -> staging for k_route: the tag is not an event flag
__global__ void k_hist_stage(View v, const Ev* in, u32 n) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Ev e = load_ev(in + i);
    if (e.dst >= v.n_lp || key_time(e.key) >= v.finish_time) { atomicOr(&v.ctl->error, E_RANGE); e.key = KEY_MAX; }
    e.flags = 0;
    store_ev(v.outbox + i, e);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) v.ctl->load_count = n;
}

// after the history replay: back to "nothing has run yet", with the history in place as processed events
__global__ void k_after_replay(View v) {
  Ctl* c = v.ctl;
  c->gvt = 0; c->local_min = KEY_MAX; c->prev_gvt = KEY_MAX; c->stall = 0;
  c->done = 0; c->round = 0; c->n_processed = 0; c->load_count = 0;
  c->rq_head = 0; c->rq_tail = 0;
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
__global__ void k_fill_u2(uint2* p, uint2 val, size_t n) {
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
