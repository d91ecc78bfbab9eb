// ssb_kernels.cuh — the Time Warp core as CUDA kernels (sm_100a).
//
// One round (= one optimistic window step), as the body of the window-loop graph (ssb_engine.cu: build_graph):
//   [k_gvt      N > 1: GVT min-reduction + barrier through the peers' mailboxes; then k_route of the received
//               anti-messages and of the received positives]
//   k_plan      GVT (block min-reduction over the calendar when N = 1), window bounds, gather list (unconsumed tails
//               of the window's buckets), scan list (their consumed heads), commit list, loop condition of the graph
//   k_commit    (parallel branch) stream-compaction of every bucket that fell below GVT into the committed log +
//               digest (+ sent-log entries for the history store); the last block does the fossil collection
//   k_exec      one thread per EVENT of the window, in calendar order (coalesced): irregularity tests (anti-message,
//               straggler against the LP's newest processed time, duplicate key through the round's hash set),
//               handler, sent-log, outbox. This single pass is the whole turn of every plain LP.
//   [k_fix_*    IF some LP is not plain: k_hist_flag (exact-differential repeat only), k_fix_count, k_fix_segs,
//               k_fix_scatter (collects the LP's arrivals and — by a coalesced scan over the consumed part of the
//               window's buckets — the processed events its rollback undoes, emitting their anti-messages),
//               k_fix_sort (segmented bitonic sort), k_fix_turn (merge, annihilate, re-execute; one warp per LP),
//               k_route(anti)]
//   k_route     outbox -> calendar buckets / the destination GPU's mailbox (TMA-staged tiles, block-aggregated atomics)
// Reference semantics restated per kernel; rule numbers R1..R13 refer to SURVEY.md Appendix A.
#pragma once
#include "ssb_device.cuh"
#include "ssb_models.cuh"

namespace ssb {

#ifdef SSB_EMU
constexpr int kRouteThreads = 1;
constexpr int kRouteItems = 32;
constexpr u32 kCommitThreads = 1;
constexpr int kSortThreads = 1;
#else
constexpr int kRouteThreads = 256;
constexpr int kRouteItems = 4;
constexpr u32 kCommitThreads = 256;
constexpr int kSortThreads = 256;
#endif
constexpr int kRouteTile = kRouteThreads * kRouteItems;   // records staged in shared memory per tile (32 KB)
constexpr u32 kCommitWarps = (kCommitThreads + 31) / 32;

constexpr u32 TICKET_PENDING = 0xFFFFFFFFu;   // lpw[lp].y while the fix-table entry of the LP is being written
constexpr u32 kHashProbes = 32;               // longer probe sequences fall back to the irregular turn
constexpr u32 kWarpSortMax = 128;             // segments up to this size are sorted by one warp in registers
constexpr u32 kSmemSortMax = 2048;            // ... up to this size by one block in shared memory, longer ones in global memory

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
  for (int o = kWarp / 2; o > 0; o >>= 1) {
    u64 w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int o = kWarp / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u64 warp_xor_u64(u64 v) {
#pragma unroll
  for (int o = kWarp / 2; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u32 warp_sum_u32(u32 v) {
#pragma unroll
  for (int o = kWarp / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u32 warp_min_u32(u32 v) {
#pragma unroll
  for (int o = kWarp / 2; o > 0; o >>= 1) { u32 w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  return v;
}
__device__ __forceinline__ u32 warp_max_u32(u32 v) {
#pragma unroll
  for (int o = kWarp / 2; o > 0; o >>= 1) { u32 w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}

// ------------------------------------------------------------------------------------------------
// GVT candidate (R7): the smallest key any unprocessed event of this partition can have. After a round every
// integrated event has been executed, so the unprocessed events are exactly the unconsumed tails of the calendar
// buckets; the start time of the first bucket with such a tail is a valid lower bound ("any value <= every LP's
// local time") and gives the same commit frontier as the exact minimum, because commit works on whole
// buckets. Events dropped beyond finish_time count as (finish_time, 0). Warp-shuffle + block min-reduction.
// Exact-differential repeat: once the recorded history of some LP is invalid from a key on, the recorded events
// behind the window count as unprocessed too (they are re-executed when the window reaches them), so the windows
// then step through the history bucket by bucket.
// Replaces scheduler::min_locals (process_scheduler.hpp:83-90) + the local half of mpi_gsync::check_sync
// (global_sync.hpp:95-150).
// ------------------------------------------------------------------------------------------------
constexpr u32 kPlanBuckets = 2048;      // calendars up to this size are staged in shared memory by k_plan
// s_fill / s_cons: optional shared-memory copies of the calendar counters (null = not staged)
__device__ __forceinline__ void block_local_min(const View& v, Ctl* c, u64* sm, u32* s_fill = nullptr, u32* s_cons = nullptr) {
  u64 m = KEY_MAX;
  for (u32 b = threadIdx.x; b < v.nb; b += blockDim.x) {
    const u32 f = v.fill[b], cs = v.consumed[b];
    if (s_fill) { s_fill[b] = f; s_cons[b] = cs; }
    if (f != cs) { u64 k = make_key(b * v.bucket_ticks, 0); m = k < m ? k : m; }
  }
  m = warp_min_u64(m);
  if ((threadIdx.x & (kWarp - 1)) == 0) sm[threadIdx.x / kWarp] = m;
  __syncthreads();
  if (threadIdx.x < kWarp) {
    m = threadIdx.x < (blockDim.x + kWarp - 1) / kWarp ? sm[threadIdx.x] : KEY_MAX;
    m = warp_min_u64(m);
    if (threadIdx.x == 0) {
      u64 bm = c->n_beyond ? make_key(v.finish_time, 0) : KEY_MAX;
      m = bm < m ? bm : m;
      if (c->n_touched && c->bound_abs < c->hist_end) {
        const u64 hm = make_key(c->bound_abs * v.bucket_ticks, 0);
        m = hm < m ? hm : m;
      }
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
  const u32 par = (c->xround + 1u) & 1u;
  const u32 R = c->xround + 1u + (c->epoch << 20);      // stamp of the exchange round that is about to start (2^20 rounds per run, 4096 resets)
  const u64 lm = c->local_min;
  const u32 my_t = lm == KEY_MAX ? 0xFFFFFFFFu : key_time(lm);
#ifdef SSB_EMU
  // host emulation (one thread per kernel; the partitions are engines of one process, a host thread each —
  // tools/emu/multi.py): the same protocol with the per-partition threads unrolled — publish to every partition
  // first, then wait for each, so that no two partitions wait for each other's publication
  for (u32 r = 0; r < v.world; ++r) {
    if (r == v.rank) continue;
    const u32 spar = c->xround & 1u;
    *reinterpret_cast<volatile u32*>(mb_cnt(v, r, spar, v.rank, 0)) = c->send_count[r];
    *reinterpret_cast<volatile u32*>(mb_cnt(v, r, spar, v.rank, 1)) = c->send_anti[r];
    __threadfence_system();
    *reinterpret_cast<volatile u64*>(mb_gvt(v, r, par, v.rank)) = ((u64)R << 32) | my_t;
  }
  for (u32 r = 0; r < v.world; ++r) {
    u32 t = my_t;
    if (r != v.rank) {
      const volatile u64* mine = mb_gvt(v, v.rank, par, r);
      u64 x = *mine;
      for (u32 spin = 0; (u32)(x >> 32) != R && spin < (1u << 27) && !(g_emu_peer_failed.load() && spin > 4096u); ++spin) { __nanosleep(20); x = *mine; }
      if ((u32)(x >> 32) != R) { atomicOr(&c->error, E_COMM); t = 0xFFFFFFFFu; }
      else t = (u32)x;
      __threadfence_system();
    }
    s_t[r] = t;
  }
#else
  if (threadIdx.x < v.world) {
    const u32 r = threadIdx.x;
    u32 t = my_t;
    if (r != v.rank) {
      // how many records this partition appended to r's mailbox since the last synchronisation (both kinds), ahead of
      // the GVT slot: when r sees the slot, the counts and the records are there
      const u32 spar = c->xround & 1u;
      *reinterpret_cast<volatile u32*>(mb_cnt(v, r, spar, v.rank, 0)) = c->send_count[r];
      *reinterpret_cast<volatile u32*>(mb_cnt(v, r, spar, v.rank, 1)) = c->send_anti[r];
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
#endif
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
// [bucket(GVT), bucket(GVT) + window); scan list = the consumed heads of the buckets a rollback of this round can
// reach (everything processed lies below the largest window bound so far). The control block is staged in shared
// memory (one coalesced load, one coalesced store): the serial bookkeeping of thread 0 then costs no global round
// trips. Nothing else touches the control block while k_plan runs (it is the first kernel of a round).
// ------------------------------------------------------------------------------------------------
__global__ void k_plan(View v, int use_global_gvt) {
  __shared__ u64 sm[32];
  __shared__ __align__(16) u32 s_words[(sizeof(Ctl) + 3) / 4];
  __shared__ u32 s_fill[kPlanBuckets], s_cons[kPlanBuckets];
  Ctl* c = reinterpret_cast<Ctl*>(s_words);
  constexpr u32 kWords = (u32)(sizeof(Ctl) / 4);
  static_assert(sizeof(Ctl) % 4 == 0, "Ctl is copied word-wise");
  for (u32 w = threadIdx.x; w < kWords; w += blockDim.x) s_words[w] = reinterpret_cast<const u32*>(v.ctl)[w];
  // the calendar counters come into shared memory with one coalesced pass, so that the serial bookkeeping below does
  // not pay a global round trip per bucket it looks at
  const bool staged = v.nb <= kPlanBuckets;
  __syncthreads();
  if (!use_global_gvt) {     // single partition: the min-reduction is fused here
    block_local_min(v, c, sm, staged ? s_fill : nullptr, staged ? s_cons : nullptr);
  } else if (staged) {
    for (u32 b = threadIdx.x; b < v.nb; b += blockDim.x) { s_fill[b] = v.fill[b]; s_cons[b] = v.consumed[b]; }
  }
  __syncthreads();
  const u32* fillp = staged ? s_fill : v.fill;
  const u32* consp = staged ? s_cons : v.consumed;
  if (threadIdx.x == 0) {
    u64 gvt = use_global_gvt ? c->gvt : c->local_min;
    c->gvt = gvt;
    u32 t = gvt == KEY_MAX ? 0xFFFFFFFFu : key_time(gvt);
    bool done = t >= v.finish_time;
    // A device error (pool / batch / ring overflow, value out of range, ...) ends the run: the host reports it after
    // the next synchronisation, and until then the rounds that are already enqueued (the unrolled body of the graph
    // loop, rounds_per_sync on the stream path) commit, gather and scan nothing — the calendar may be inconsistent.
    const bool halted = c->error != 0u;
    u32 fin_b = (v.finish_time + v.bucket_ticks - 1) / v.bucket_ticks;   // buckets [0, fin_b) hold times < finish
    if (fin_b > v.nb) fin_b = v.nb;
    u32 cur = done ? fin_b : t / v.bucket_ticks;
    if (cur > fin_b) cur = fin_b;
    c->cur_abs = cur;
    // commit list (a recorded history that was never re-executed is not committed: k_commit skips P_HISTORY)
    u32 nc = 0, tot = 0;
    for (u32 b = c->commit_next; b < cur && !halted; ++b) {
      u32 f = fillp[b];
      if (f) { v.commit[nc++] = make_uint4(b, f, tot, 0); tot += f; }
    }
    if (!halted) c->commit_next = cur > c->commit_next ? cur : c->commit_next;
    c->n_commit = nc;
    c->commit_total = tot;
    // snapshot ring: rounds whose window lies below GVT are dead (their snapshots can be reused)
    while (c->rq_tail != c->rq_head && c->rq_bound[c->rq_tail % RQ_LEN] <= cur) c->rq_tail++;
    c->snap_tail = c->rq_tail == c->rq_head ? c->snap_head : c->rq_snap[c->rq_tail % RQ_LEN];
    u32 ng = 0, total = 0, ns = 0, stot = 0;
    u32 bound = cur;
    if (!done && !halted) {
      // scan list: what has been processed in the buckets from the window start on (before this round's gather)
      u32 shi = cur + v.window < fin_b ? cur + v.window : fin_b;
      if (c->hi_bound > shi) shi = c->hi_bound < fin_b ? c->hi_bound : fin_b;
      for (u32 b = cur; b < shi; ++b) {
        const u32 cons = consp[b];
        if (cons) { v.scan[ns++] = make_uint4(b, 0u, cons, stot); stot += cons; }
      }
      // gather list
      bound = cur + v.window < fin_b ? cur + v.window : fin_b;
      if (v.sync > 1) {
        // several windows per global synchronisation: GVT lags behind the partition's progress (it only moves at a
        // synchronisation), so a window is [GVT bucket, progress frontier + 1): whatever arrived late below the
        // frontier (stragglers) plus ONE new bucket — unless the partition is already kLeadMax buckets ahead of GVT
        const u32 kLeadMax = 4u * v.sync;
        u32 nb2 = c->hi_bound + 1u;
        if (c->hi_bound >= cur + kLeadMax) nb2 = c->hi_bound;
        if (nb2 > bound) bound = nb2 < fin_b ? nb2 : fin_b;
      }
      for (u32 b = cur; b < bound; ++b) {
        const u32 cons = consp[b];
        u32 avail = fillp[b] - cons;
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
      if (bound > c->hi_bound) c->hi_bound = bound;
    }
    c->bound_abs = bound;
    c->n_gather = ng;
    c->n_arr = total;
    {      // hash-set region of this round: the smallest power of two >= 2 * arrivals (at least 1024 slots)
      // at least 2 x the arrivals; 4 x while the region stays below 8 MB (fewer second probes, still L2-resident)
      u32 hm = 1023u;
      while (hm < v.hmask && (hm + 1u < 2u * total || (hm + 1u < v.hash_hi * total && hm + 1u < v.hash_cap))) hm = (hm << 1) | 1u;
      c->hmask_round = hm < v.hmask ? hm : v.hmask;
    }
    c->n_scan = ns;
    c->scan_total = stot;
    c->n_processed += total;   // k_fix_turn corrects this for the LPs it re-executes
    c->n_anti = 0;
    c->n_extra = 0;
    c->n_fix = 0;
    c->n_ent = 0;
    c->fix_min_t = 0xFFFFFFFFu;
    c->done = done ? 1u : 0u;
    c->log_stable = c->log_count;      // the commit passes of all earlier rounds are complete
    c->round++;
    c->stamp_base += 2u;
    if (v.sub == 0) {
      // first window after a global synchronisation: a new exchange round starts (what this partition sends until the
      // next synchronisation accumulates in one mailbox half and in sent_min)
      c->xround++;
      for (int i = 0; i < MAX_WORLD; ++i) { c->send_count[i] = 0; c->send_anti[i] = 0; }
      c->sent_min = KEY_MAX;
    }
    if (v.use_graph) {
      // GVT must keep advancing (every round commits or integrates something): give up instead of spinning forever
      if (v.sub == 0) { if (gvt == c->prev_gvt) { if (++c->stall > kStallRounds) c->error |= E_STALL; } else { c->stall = 0; c->prev_gvt = gvt; } }
      cudaGraphSetConditional(v.h_loop, (done || c->error) ? 0u : 1u);
    }
    // (the mailbox half that was just integrated needs no clearing: its counts are overwritten by the senders before
    // the half is read again)
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
__global__ void __launch_bounds__(kCommitThreads) k_commit(View v, typename M::Data md) {
  Ctl* c = v.ctl;
  const u32 total = c->commit_total;
  const u32 nc = c->n_commit;
  if (nc == 0) return;
  u64 cnt = 0, hsum = 0, hxor = 0, hsum2 = 0;
  const u32 lane = threadIdx.x & (kWarp - 1), warp = threadIdx.x / kWarp;
  __shared__ u32 w_cnt[kCommitWarps];
  __shared__ u64 s_pos;
  __shared__ u64 s_dg[kCommitWarps][4];
  __shared__ u32 s_last;
  // every thread of a block iterates the same number of times (block-level compaction inside the loop)
  for (u32 base = blockIdx.x * kCommitThreads; base < total; base += gridDim.x * kCommitThreads) {
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
        for (u32 w = 0; w < kCommitWarps; ++w) { u32 a = w_cnt[w]; w_cnt[w] = tot; tot += a; }
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
    for (u32 w = 0; w < kCommitWarps; ++w) { a += s_dg[w][0]; b += s_dg[w][1]; x ^= s_dg[w][2]; d += s_dg[w][3]; }
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
    if (nch > v.maxc) nch = v.maxc;      // (only after an overflow; k_plan commits nothing once an error is set)
    for (u32 j = threadIdx.x; j < nch; j += blockDim.x) {
      u32* entry = &v.chunk_tab[(size_t)b * v.maxc + j];
      u32 ch = *entry;
      *entry = NIL;
      if (ch == NIL || ch == 0u) continue;
      u32 pos = atomicAdd(&c->free_top, 1u);
      v.free_stack[pos] = ch;
    }
    if (threadIdx.x == 0) { v.fill[b] = 0; v.consumed[b] = 0; }
  }
  if (threadIdx.x == 0) c->commit_blocks = 0;
}

// ------------------------------------------------------------------------------------------------
// calendar slot of the arrival with index i (arrival order = calendar order of the gathered bucket tails), and of
// item i of the scan list (the consumed heads)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 slot_of_list(const View& v, const uint4* list, u32 n, u32 i) {
  u32 g = 0, hi = n;
  while (hi - g > 1) {
    u32 mid = (g + hi) >> 1;
    if (list[mid].w <= i) g = mid; else hi = mid;
  }
  const uint4 ge = list[g];
  return slot_of(v, ge.x, ge.y + (i - ge.w));
}
__device__ __forceinline__ u32 slot_of_arrival(const View& v, u32 i) { return slot_of_list(v, v.gather, v.ctl->n_gather, i); }

// ------------------------------------------------------------------------------------------------
// Irregular turns. An LP whose turn of this round is not plain gets a ticket (index into the fix table); the first
// flagger issues it. lpw[lp].y: 0 = plain, TICKET_PENDING while the table entry is written, then ticket + 1.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void flag_lp(const View& v, u32 lp, u32 time) {
  atomicMin(&v.ctl->fix_min_t, time);
  u32* t = &v.lpw[lp].y;
  if (atomicCAS(t, 0u, TICKET_PENDING) != 0u) return;
  const u32 f = atomicAdd(&v.ctl->n_fix, 1u);      // < n_lp_local: every LP is issued at most one ticket per round
  v.fixlp[f] = lp;
  v.fixcnt[f] = 0;
  v.fixund[f] = 0;
  v.fixmin[f] = KEY_MAX;
  __threadfence();
  atomicExch(t, f + 1u);
}

// Duplicate detection (R1: "same key twice at one LP = one event", queue.hpp:92) without per-LP lists: every plain
// arrival inserts a 32-bit fingerprint of (LP, key) into the round's hash set (open addressing, linear probing, ONE
// atomic per probe). The table region used by a round is sized by the round's arrivals (Ctl::hmask_round), so it stays
// L2-resident; k_post clears it again behind k_exec, off the critical path.
// Returns false when the same fingerprint is already there (duplicate SUSPECT: the LP takes the exact irregular
// turn) or the probe sequence gets too long. Two arrivals with the same (LP, key) follow the same probe sequence,
// and occupied slots stay occupied during a round, so at most one of them can return true.
__device__ __forceinline__ u64 hash_of(u32 lp, u64 key) { return mix64(key ^ ((u64)lp * 0x9E3779B97F4A7C15ULL)); }
// probes from `pos` on (fingerprint fp, never 0 = free slot)
__device__ __forceinline__ bool hash_insert_from(const View& v, u32 fp, u32 pos, u32 hmask) {
  for (u32 probe = 1; probe < kHashProbes; ++probe) {
    const u32 old = atomicCAS(v.hset + pos, 0u, fp);
    if (old == 0u) return true;
    if (old == fp) return false;
    pos = (pos + 1u) & hmask;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
// k_exec — one thread per EVENT of the window, in calendar (slot) order, so the event read and the meta / pflag /
// outbox writes are coalesced streams; only the per-LP word, the hash-set slot and the model's tables are gathered.
// This is the speculative form of runner::run_lp_aggr (runner.hpp:530-570): every event of an LP is executed against
// the LP's current state, which is exactly the sequential result as long as no invocation changes the state (neither
// shipped application ever does) and the LP's turn is plain:
//   * no anti-message among its arrivals                                   (R4/R6, queue.hpp:87-90)
//   * no arrival at or before the time of its newest processed event       (straggler, R5, queue.hpp:98-105);
//     lpw[lp].x is stable during the pass — k_route raises it afterwards, from the send time of every output
//   * no key twice                                                          (R1, queue.hpp:92)
// Otherwise the LP is flagged and k_fix_* redo its whole turn exactly (whatever this pass wrote for the LP's other
// arrivals is overwritten: arrival i owns outbox entry i). Skew-free: a hot LP's events are spread over many threads.
// ------------------------------------------------------------------------------------------------
// LP state -> registers with the widest loads the record size allows (8 words = two 128-bit loads of one sector)
template <int W>
__device__ __forceinline__ void load_state(const u32* p, u32 (&st)[W]) {
  if (W % 4 == 0) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) { const uint4 x = reinterpret_cast<const uint4*>(p)[q]; st[4 * q] = x.x; st[4 * q + 1] = x.y; st[4 * q + 2] = x.z; st[4 * q + 3] = x.w; }
  } else if (W % 2 == 0) {
#pragma unroll
    for (int q = 0; q < W / 2; ++q) { const uint2 x = reinterpret_cast<const uint2*>(p)[q]; st[2 * q] = x.x; st[2 * q + 1] = x.y; }
  } else {
#pragma unroll
    for (int q = 0; q < W; ++q) st[q] = p[q];
  }
}

// handler on state words that are already in registers; returns true when the invocation changed the state
template <class M>
__device__ __forceinline__ bool run_handler(const typename M::Data& md, const Ev& e, u32 (&st)[M::kStateWords], u32 early,
                                            Ev& out, bool& has, u32& err) {
  u32 pre[M::kStateWords];
#pragma unroll
  for (int q = 0; q < M::kStateWords; ++q) pre[q] = st[q];
  Ev in = e;
  in.flags = 0;
  has = M::handle(md, in, st, early, out, err);
  bool changed = false;
#pragma unroll
  for (int q = 0; q < M::kStateWords; ++q) changed |= (pre[q] != st[q]);
  return has && changed;
}

// Measured on B200 (PHOLD config 2, round 2): 1 / 2 / 4 events in flight per thread give 7.07 / 7.08 / 8.9 ms per
// simulation — the extra registers cost the occupancy that the extra loads would have used. One event per thread.
#ifndef SSB_EXEC_UNROLL
#define SSB_EXEC_UNROLL 1
#endif
constexpr int kExecUnroll = SSB_EXEC_UNROLL;      // events in flight per thread: the pass is bound by memory latency, not by issue slots

// An arrival whose first hash probe hits a slot that another key occupies is executed like a plain one, and the rest of
// its probe sequence is parked in shared memory: the block walks all parked sequences together at the end, so that the
// dependent atomics of a collision cost one round trip chain per BLOCK instead of one per warp and loop iteration
// (measured: the in-loop probes were a third of the kernel's stall samples). If the sequence ends in a duplicate
// suspect, the LP takes its ticket then — k_fix_turn redoes the whole turn and overwrites what this pass wrote.
constexpr u32 kExecParked = 512;

template <class M>
__global__ void __launch_bounds__(256, M::kExecMinBlocks) k_exec(View v, typename M::Data md) {
  __shared__ uint4 s_parked[kExecParked];      // {next probe position, fingerprint, LP, receive time}
  __shared__ u32 s_nparked;
  Ctl* c = v.ctl;
  const u32 n_arr = c->n_arr;
  const u32 hmask = c->hmask_round;
  const u32 nth = gridDim.x * blockDim.x;
  u32 err = 0;
  if (threadIdx.x == 0) s_nparked = 0;
  __syncthreads();
  for (u32 i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n_arr; i0 += kExecUnroll * nth) {
    // Per event two dependent levels of memory latency instead of a chain: (1) the event record; (2) everything the
    // turn needs, issued together — the per-LP word, the LP state, the model's gather (M::early) and the first probe
    // of the hash set (inserting the fingerprint of an arrival that turns out to be irregular is harmless: its LP takes
    // a ticket anyway); then decisions, handler arithmetic and the coalesced stores. kExecUnroll events per thread.
    u32 slot[kExecUnroll], lp[kExecUnroll], hpos[kExecUnroll], hfp[kExecUnroll], hold[kExecUnroll], early[kExecUnroll];
    u32 st[kExecUnroll][M::kStateWords];
    Ev e[kExecUnroll];
    uint2 w[kExecUnroll];
    bool live[kExecUnroll];
#pragma unroll
    for (int u = 0; u < kExecUnroll; ++u) {
      const u32 i = i0 + (u32)u * nth;
      live[u] = i < n_arr;
      // (evict-first: the record is not read again before its bucket is committed; measured 1 % on the whole loop.
      // An evict-LAST policy on the gathered tables — per-LP word, state, lookup table — changed nothing: they stay in
      // L2 as it is)
      if (live[u]) { slot[u] = slot_of_arrival(v, i); e[u] = load_ev_ef(v.ev + slot[u]); }
    }
#pragma unroll
    for (int u = 0; u < kExecUnroll; ++u) {
      if (!live[u]) continue;
      lp[u] = part_local(v, e[u].dst);
      w[u] = __ldcg(reinterpret_cast<const uint2*>(v.lpw + lp[u]));
      if (M::kUsesState) load_state<M::kStateWords>(v.state + (size_t)lp[u] * M::kStateWords, st[u]);
      else {
#pragma unroll
        for (int q = 0; q < M::kStateWords; ++q) st[u][q] = 0u;
      }
      early[u] = M::early(md, e[u]);
      hold[u] = 0u;
      if (!(e[u].flags & F_CANCEL)) {
        const u64 h = hash_of(lp[u], e[u].key);
        hfp[u] = (u32)(h >> 32) | 1u;
        hpos[u] = (u32)h & hmask;
        hold[u] = atomicCAS(v.hset + hpos[u], 0u, hfp[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kExecUnroll; ++u) {
      if (!live[u]) continue;
      const u32 i = i0 + (u32)u * nth;
      bool bad = (e[u].flags & F_CANCEL) != 0 || key_time(e[u].key) < w[u].x;
      if (w[u].y) {
        // the LP already has a ticket: k_fix_turn runs its turn. Every arrival that can undo something still lowers
        // the bucket from which the rollback scans start
        if (bad) atomicMin(&c->fix_min_t, key_time(e[u].key));
        continue;
      }
      if (!bad && hold[u] != 0u) {
        if (hold[u] == hfp[u]) bad = true;
        else {
          const u32 d = atomicAdd(&s_nparked, 1u);
          if (d < kExecParked) s_parked[d] = make_uint4((hpos[u] + 1u) & hmask, hfp[u], lp[u], key_time(e[u].key));
          else bad = !hash_insert_from(v, hfp[u], (hpos[u] + 1u) & hmask, hmask);
        }
      }
      if (bad) { flag_lp(v, lp[u], key_time(e[u].key)); continue; }
      Ev out;
      bool has;
      if (run_handler<M>(md, e[u], st[u], early[u], out, has, err)) { flag_lp(v, lp[u], key_time(e[u].key)); continue; }      // state-mutating handler
      Meta m;
      m.sent_dst = NIL; m.snap = NIL; m.sent_key = 0;
      if (has) {
        m.sent_dst = out.dst;
        m.sent_key = out.key;
        store_ev(v.outbox + i, out);
      } else {
        v.outbox[i].key = KEY_MAX;
      }
      store_meta_ef(v.meta + slot[u], m);
      v.pflag[slot[u]] = P_PROCESSED;
    }
  }
  __syncthreads();
  {
    const u32 np = s_nparked < kExecParked ? s_nparked : kExecParked;
    for (u32 j = threadIdx.x; j < np; j += blockDim.x) {
      const uint4 q = s_parked[j];
      if (!hash_insert_from(v, q.y, q.x, hmask)) flag_lp(v, q.z, q.w);
    }
  }
  if (err) atomicOr(&c->error, err);
  if (v.use_graph) {
    // the last block to finish knows whether any LP was flagged: it opens or closes the IF node around k_fix_*
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&c->exec_blocks, 1u) == gridDim.x - 1) {
        c->exec_blocks = 0;
        const u32 nf = *reinterpret_cast<volatile u32*>(&c->n_fix);
        cudaGraphSetConditional(v.h_fix, (nf || (v.inv && c->n_touched)) ? 1u : 0u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_post — behind k_exec, off the critical path (in the graph it runs next to k_route): the hash-set region the round
// used is cleared again. (The other piece of bookkeeping after a pass — the output of arrival i raises its sender's
// newest processed time lpw[src].x, which the straggler test of the next round's k_exec reads — is done by k_route,
// which has the outputs staged in shared memory anyway: a separate pass over the outbox cost 46 MB of reads per
// round.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_post(View v) {
  Ctl* c = v.ctl;
  if (!c->n_arr) return;
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  uint4* hs = reinterpret_cast<uint4*>(v.hset);
  const u32 nq = (c->hmask_round + 1u) >> 2;
  for (u32 i = tid; i < nq; i += nth) hs[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------------
// The irregular turns, k_fix_*: what eventq::merge_buffer + lp::flush_buf + the re-execution in run_lp_aggr do for an
// LP with an anti-message, a duplicate, a straggler or a state-mutating handler (queue.hpp:82-108,
// logical_process.hpp:129-157, runner.hpp:530-570) — as data-parallel passes:
//   k_hist_flag    exact-differential repeat only: a recorded event inside the window whose LP's history is invalid
//                  from an earlier key on must be re-executed: its LP takes a ticket                          (R13)
//   k_fix_count    per ticket: number of entries (its arrivals + every processed event of the scan list that belongs
//                  to the LP: an upper bound of what the rollback undoes) and the smallest arrival key m    (R4)
//   k_fix_segs     one segment of `sorted` per ticket
//   k_fix_scatter  arrivals -> entries; scan list: every processed event of a ticketed LP with key >= m is UNDONE —
//                  its recorded output leaves as an anti-message, its processed flag is cleared, it becomes an entry
//                  of the segment (it was in the queue before every arrival)                                   (R5)
//   k_fix_sort     segments sorted by (key, arrival order)                                                     (R1)
//   k_fix_turn     one warp per LP: per key apply the inbox items in arrival order (positive = insert unless
//                  present, anti = erase if present), restore the state snapshot, execute the survivors        (R2, R4, R6)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 ticket_of(const View& v, u32 lp) {
  const u32 t = *reinterpret_cast<volatile u32*>(&v.lpw[lp].y);
  return t;      // 0 = none; TICKET_PENDING cannot be seen after the kernel that issued it
}

__global__ void __launch_bounds__(256) k_hist_flag(View v) {
  Ctl* c = v.ctl;
  if (!v.inv || !c->n_touched) return;
  const u32 n = c->scan_total;
  const u32 tbound = c->bound_abs * v.bucket_ticks;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of_list(v, v.scan, c->n_scan, i);
    const unsigned char pf = v.pflag[slot];
    if ((pf & (P_PROCESSED | P_HISTORY | P_DEAD)) != (P_PROCESSED | P_HISTORY)) continue;
    const Ev e = load_ev(v.ev + slot);
    if (key_time(e.key) >= tbound) continue;
    const u32 lp = part_local(v, e.dst);
    if (e.key >= v.inv[lp]) flag_lp(v, lp, key_time(e.key));
  }
}

// first item of the scan list that a rollback of this round can reach: the scan entries are in bucket order, and
// nothing below the bucket of the earliest flagging arrival is undone
__device__ __forceinline__ u32 scan_begin(const View& v, const Ctl* c) {
  const u32 bmin = c->fix_min_t / v.bucket_ticks;
  u32 off = c->scan_total;
  for (u32 k = 0; k < c->n_scan; ++k) {
    const uint4 se = v.scan[k];
    if (se.x >= bmin) { off = se.w; break; }
  }
  return off;
}

__global__ void __launch_bounds__(256) k_fix_count(View v) {
  Ctl* c = v.ctl;
  if (c->n_fix == 0) return;
  const u32 s0 = scan_begin(v, c);
  const u32 n_arr = c->n_arr, n = n_arr + c->scan_total - s0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < n_arr) {
      const u32 slot = slot_of_arrival(v, i);
      const Ev e = load_ev(v.ev + slot);
      const u32 t = ticket_of(v, part_local(v, e.dst));
      if (!t) continue;
      atomicAdd(&v.fixcnt[t - 1u], 1u);
      atomicMin(&v.fixmin[t - 1u], e.key);
    } else {
      const u32 slot = slot_of_list(v, v.scan, c->n_scan, s0 + (i - n_arr));
      const unsigned char pf = v.pflag[slot];
      if ((pf & (P_PROCESSED | P_DEAD)) != P_PROCESSED) continue;
      const Ev e = load_ev(v.ev + slot);
      const u32 t = ticket_of(v, part_local(v, e.dst));
      if (t) atomicAdd(&v.fixcnt[t - 1u], 1u);
    }
  }
}

__global__ void k_fix_segs(View v) {
  Ctl* c = v.ctl;
  const u32 nf = c->n_fix;
  for (u32 f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    u32 n = v.fixcnt[f];
    u32 start = atomicAdd(&c->n_ent, n);
    if (start + n > v.cap_fix || start + n < start) { atomicOr(&c->error, E_BATCH); start = 0; n = 0; v.fixcnt[f] = 0; }
    v.fixstart[f] = start;
    v.fixpos[f] = start;
  }
}

__device__ __forceinline__ void store_sorted(Sorted* p, u64 key, u32 slot, u32 seq) {
  *reinterpret_cast<uint4*>(p) = make_uint4((u32)key, (u32)(key >> 32), slot, seq);
}
__device__ __forceinline__ Sorted load_sorted(const Sorted* p) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  Sorted x;
  x.key = ((u64)q.y << 32) | q.x; x.slot = q.z; x.seq = q.w;
  return x;
}
// claims the next entry of ticket t's segment
__device__ __forceinline__ Sorted* claim_entry(const View& v, u32 t) {
  const u32 f = t - 1u;
  const u32 p = atomicAdd(&v.fixpos[f], 1u);
  if (p >= v.fixstart[f] + v.fixcnt[f]) {
    if (!(v.ctl->error & E_BATCH)) internal_error(v, 6, v.fixlp[f]);      // (a segment voided by E_BATCH takes no entries)
    return nullptr;
  }
  return v.sorted + p;
}

// entries of ticket f's segment. (A segment that did not fit the entry table — E_BATCH, k_fix_segs — has fixcnt 0
// while claim_entry still counted the attempts in fixpos: it is empty, not `attempts` entries of someone else's data.)
__device__ __forceinline__ u32 seg_len(const View& v, u32 f) {
  const u32 claimed = v.fixpos[f] - v.fixstart[f], counted = v.fixcnt[f];
  return claimed < counted ? claimed : counted;
}

__global__ void __launch_bounds__(256) k_fix_scatter(View v) {
  Ctl* c = v.ctl;
  if (c->n_fix == 0) return;
  const u32 s0 = scan_begin(v, c);
  const u32 n_arr = c->n_arr, n = n_arr + c->scan_total - s0;
  const u32 tbound = c->bound_abs * v.bucket_ticks;
  u32 n_und = 0, n_anti = 0, err = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < n_arr) {
      const u32 slot = slot_of_arrival(v, i);
      const Ev e = load_ev(v.ev + slot);
      const u32 t = ticket_of(v, part_local(v, e.dst));
      if (!t) continue;
      Sorted* p = claim_entry(v, t);
      if (p) store_sorted(p, e.key, slot, (i + 1u) | ((e.flags & F_CANCEL) ? S_ANTI : 0u));
    } else {
      const u32 slot = slot_of_list(v, v.scan, c->n_scan, s0 + (i - n_arr));
      const unsigned char pf = v.pflag[slot];
      if ((pf & (P_PROCESSED | P_DEAD)) != P_PROCESSED) continue;
      const Ev e = load_ev(v.ev + slot);
      const u32 lp = part_local(v, e.dst);
      const u32 t = ticket_of(v, lp);
      if (!t) continue;
      // R5: undone = every processed event of the LP with key >= m. A recorded event behind the window is left in
      // place (its LP's inv key marks it invalid; it is undone when the window reaches it).
      const bool hist = (pf & P_HISTORY) != 0;
      if (hist && key_time(e.key) >= tbound) continue;
      if (!(e.key >= v.fixmin[t - 1u] || (hist && v.inv && e.key >= v.inv[lp]))) continue;
      const Meta mt = load_meta(v.meta + slot);
      if (mt.sent_dst != NIL) {      // the sent-log entry leaves as an anti-message (queue.hpp:99-105)
        const u32 pos = atomicAdd(&c->n_anti, 1u);
        if (pos >= v.cap_aux) err |= E_BATCH;
        else {
          Ev a;
          a.key = mt.sent_key; a.dst = mt.sent_dst; a.src = e.dst; a.send = key_time(e.key); a.p0 = 0; a.p1 = 0; a.flags = F_CANCEL;
          store_ev(v.anti_box + pos, a);
          ++n_anti;
        }
      }
      v.pflag[slot] = 0;
      atomicAdd(&v.fixund[t - 1u], 1u);
      ++n_und;
      Sorted* p = claim_entry(v, t);
      if (p) store_sorted(p, e.key, slot, S_UNDONE | (mt.snap != NIL ? (mt.snap + 1u) & S_IDX : 0u));
    }
  }
  n_und = warp_sum_u32(n_und);
  n_anti = warp_sum_u32(n_anti);
  if ((threadIdx.x & (kWarp - 1)) == 0) {
    if (n_und) atomicAdd(&c->n_undone, (u64)n_und);
    if (n_anti) atomicAdd(&c->n_antis, (u64)n_anti);
  }
  if (err) atomicOr(&c->error, err);
}

// ------------------------------------------------------------------------------------------------
// k_fix_sort — every segment sorted by (key, arrival order) (R1; an undone event precedes the arrivals under its key:
// it was in the queue before them).
//   n <= kWarpSortMax : one warp, bitonic network over register shuffles, 1 / 2 / 4 entries per lane
//   n <= kSmemSortMax : one block, bitonic network in shared memory (all compare-exchanges ascending, so the
//                       padding up to the next power of two needs no storage: pairs that reach beyond n are skipped)
//   longer            : the same network in global memory
// ------------------------------------------------------------------------------------------------
struct KS { u64 key; u32 slot, seq; };
__device__ __forceinline__ bool ks_less(const KS& a, const KS& b) {
  return a.key < b.key || (a.key == b.key && sorted_ord(a.seq) < sorted_ord(b.seq));
}
__device__ __forceinline__ KS ks_load(const uint4* p) {
  const uint4 q = *p;
  KS x;
  x.key = ((u64)q.y << 32) | q.x; x.slot = q.z; x.seq = q.w;
  return x;
}
__device__ __forceinline__ void ks_store(uint4* p, const KS& x) { *p = make_uint4((u32)x.key, (u32)(x.key >> 32), x.slot, x.seq); }

#ifndef SSB_EMU
__device__ __forceinline__ KS ks_shfl_xor(const KS& x, int m) {
  KS y;
  const u32 lo = __shfl_xor_sync(0xffffffffu, (u32)x.key, m), hi = __shfl_xor_sync(0xffffffffu, (u32)(x.key >> 32), m);
  y.key = ((u64)hi << 32) | lo;
  y.slot = __shfl_xor_sync(0xffffffffu, x.slot, m);
  y.seq = __shfl_xor_sync(0xffffffffu, x.seq, m);
  return y;
}
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
__device__ __forceinline__ void warp_sort_segment(uint4* seg, u32 n, u32 lane) {
  KS x[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u32 i = lane + ((u32)r << 5);
    if (i < n) x[r] = ks_load(seg + i);
    else { x[r].key = KEY_MAX; x[r].slot = NIL; x[r].seq = S_IDX; }
  }
  warp_bitonic<R>(x, lane);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const u32 i = lane + ((u32)r << 5);
    if (i < n) ks_store(seg + i, x[r]);
  }
}
#endif

// block-wide bitonic network over a[0..n) (shared or global memory), every compare-exchange ascending: the first
// step of a merge level pairs i with its mirror inside the 2^k block, the following steps pair i with i ^ j
__device__ __forceinline__ void block_bitonic(uint4* a, u32 n) {
  u32 P = 1;
  while (P < n) P <<= 1;
  for (u32 k = 2; k <= P; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      const u32 mask = (j == (k >> 1)) ? (k - 1u) : j;
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const u32 l = i ^ mask;
        if (l > i && l < n) {
          const KS x = ks_load(a + i), y = ks_load(a + l);
          if (ks_less(y, x)) { ks_store(a + i, y); ks_store(a + l, x); }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kSortThreads) k_fix_sort(View v) {
  __shared__ uint4 s_seg[kSmemSortMax];
  Ctl* c = v.ctl;
  const u32 nf = c->n_fix;
  if (nf == 0) return;
  uint4* sorted = reinterpret_cast<uint4*>(v.sorted);
#ifndef SSB_EMU
  // short segments: one warp each
  const u32 lane = threadIdx.x & 31u;
  const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 f = wid; f < nf; f += nw) {
    const u32 start = v.fixstart[f], n = seg_len(v, f);
    if (n < 2 || n > kWarpSortMax) continue;
    if (n <= 32) warp_sort_segment<1>(sorted + start, n, lane);
    else if (n <= 64) warp_sort_segment<2>(sorted + start, n, lane);
    else warp_sort_segment<4>(sorted + start, n, lane);
  }
  const u32 small_max = kWarpSortMax;
#else
  const u32 small_max = 1;
#endif
  // long segments: one block each
  for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
    const u32 start = v.fixstart[f], n = seg_len(v, f);
    if (n <= small_max) continue;
    if (n <= kSmemSortMax) {
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) s_seg[i] = sorted[start + i];
      __syncthreads();
      block_bitonic(s_seg, n);
      for (u32 i = threadIdx.x; i < n; i += blockDim.x) sorted[start + i] = s_seg[i];
      __syncthreads();
    } else {
      __syncthreads();
      block_bitonic(sorted + start, n);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fix_turn — one warp per irregular LP; its segment holds the arrivals of the round and the undone events, sorted.
//   A. merge (R4/R6, queue.hpp:82-96): per key, in arrival order: the undone event (if any) is present; a positive is
//      inserted unless the key is present (else it is a duplicate and dies); an anti-message erases what is present
//      (annihilation) and is consumed. Lanes work on different keys; a key group longer than one entry is walked by
//      the lane of its first entry. Result per key: the surviving event and its outbox entry.
//   B. state restore (R5, queue.hpp:286-290): the pre-state of the smallest undone event that changed the state.
//   C. execute the survivors (R2, runner.hpp:544-568): speculatively in parallel against that state — exact unless an
//      invocation changes the state; if one does, lane 0 redoes the turn sequentially with copy-on-change snapshots.
//      A survivor behind the window (an undone event of a later bucket) goes back to PENDING: a copy is appended to
//      its bucket through the outbox, the recorded slot dies — what lp::flush_buf achieves by lowering local_time_
//      (logical_process.hpp:145-156).
// ------------------------------------------------------------------------------------------------
template <class M>
struct TurnExec {
  const View& v;
  const typename M::Data& md;
  u32 err;
  __device__ TurnExec(const View& v_, const typename M::Data& md_) : v(v_), md(md_), err(0) {}

  __device__ __forceinline__ u32 extra_idx() {
    u32 pos = atomicAdd(&v.ctl->n_extra, 1u);
    if (pos >= v.cap_aux) { err |= E_BATCH; pos = v.cap_aux - 1; }
    return v.cap_batch + pos;
  }
  // handler on a copy of `st`; returns true when the invocation changed the state (nothing is written then)
  __device__ __forceinline__ bool try_process(u32 slot, u32 out_idx, const u32* st) {
    const Ev e0 = load_ev(v.ev + slot);
    Ev e = e0;
    e.flags = 0;
    u32 ls[M::kStateWords];
#pragma unroll
    for (int w = 0; w < M::kStateWords; ++w) ls[w] = st[w];
    Ev out;
    const bool has = M::handle(md, e, ls, M::early(md, e), out, err);
    if (has) {
      bool changed = false;
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) changed |= (ls[w] != st[w]);
      if (changed) return true;
    }
    Meta m;
    m.sent_dst = NIL; m.snap = NIL; m.sent_key = 0;
    if (has) { m.sent_dst = out.dst; m.sent_key = out.key; store_ev(v.outbox + out_idx, out); }
    else v.outbox[out_idx].key = KEY_MAX;
    store_meta(v.meta + slot, m);
    v.pflag[slot] = P_PROCESSED;
    return false;
  }
  // sequential form: the state evolves, the pre-state of a changing invocation goes to the snapshot ring
  __device__ __forceinline__ void process_seq(u32 slot, u32 out_idx, u32* st, bool& dirty) {
    Ev e = load_ev(v.ev + slot);
    e.flags = 0;
    u32 pre[M::kStateWords];
#pragma unroll
    for (int w = 0; w < M::kStateWords; ++w) pre[w] = st[w];
    Ev out;
    const bool has = M::handle(md, e, st, M::early(md, e), out, err);
    Meta m;
    m.sent_dst = NIL; m.snap = NIL; m.sent_key = 0;
    if (has) {
      bool changed = false;
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) changed |= (pre[w] != st[w]);
      if (changed) {
        const u32 s = atomicAdd(&v.ctl->snap_head, 1u);
        if (s - v.ctl->snap_tail >= v.cap_snap) err |= E_SNAP;
        const u32 si = s % v.cap_snap;
#pragma unroll
        for (int w = 0; w < M::kStateWords; ++w) v.snap[(size_t)si * M::kStateWords + w] = pre[w];
        m.snap = si;
        dirty = true;
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
  }
  __device__ __forceinline__ void requeue(u32 slot) {
    Ev e = load_ev(v.ev + slot);
    e.flags &= ~((1u << F_STAMP_SHIFT) - 1u);      // still the same event: it keeps its place in its sender's sequence
    store_ev(v.outbox + extra_idx(), e);
    v.pflag[slot] = P_DEAD;
  }
};

template <class M>
__global__ void __launch_bounds__(256) k_fix_turn(View v, typename M::Data md) {
  Ctl* c = v.ctl;
  const u32 nf = c->n_fix;
  if (nf == 0) return;
  const u32 lane = threadIdx.x & (kWarp - 1);
  const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = (gridDim.x * blockDim.x) / kWarp;
  const u32 tbound = c->bound_abs * v.bucket_ticks;
  TurnExec<M> T(v, md);
  u32 s_proc = 0, s_arr = 0, s_annih = 0, s_dup = 0, s_rb = 0;
  for (u32 f = wid; f < nf; f += nw) {
    const u32 lp = v.fixlp[f];
    const u32 start = v.fixstart[f], n = seg_len(v, f);
    const u64 m = v.fixmin[f];
    const u32 n_und = v.fixund[f];
    const Sorted* seg = v.sorted + start;
    uint2* eo = v.ent_out + start;
    // --- A. merge ---
    u32 snap_pos = NIL;      // first (smallest key) undone entry that carries a snapshot
    for (u32 i = lane; i < n; i += kWarp) {
      const Sorted a = load_sorted(seg + i);
      if (!(a.seq & S_UNDONE)) ++s_arr;
      else if ((a.seq & S_IDX) && i < snap_pos) snap_pos = i;
      if (i > 0 && seg[i - 1].key == a.key) { eo[i] = make_uint2(NIL, NIL); continue; }      // not the first entry of its key
      u32 present = NIL, out_idx = NIL;
      // the entries of this key: [i, jend). A lone entry (the rule) needs no order; several are applied in the order
      // their sender emitted them (sequence stamp, then arrival index), the undone event first
      u32 jend = i + 1u;
      while (jend < n && seg[jend].key == a.key) ++jend;
      u64 last = 0;                                       // (stamp, arrival index + 1) of the entry applied last
      for (u32 step = i; step < jend; ++step) {
        Sorted b = a;
        if (jend - i > 1u) {
          u64 best = ~0ull;
          for (u32 j = i; j < jend; ++j) {
            const Sorted x = load_sorted(seg + j);
            const u64 ord = (x.seq & S_UNDONE) ? 1ull
                                               : (((u64)(v.ev[x.slot].flags >> F_STAMP_SHIFT) + 1ull) << 32) | (x.seq & S_IDX);
            if (ord > last && ord < best) { best = ord; b = x; }
          }
          last = best;
        }
        if (b.seq & S_UNDONE) {
          present = b.slot;                               // was in the queue before the arrivals
        } else if (b.seq & S_ANTI) {
          v.pflag[b.slot] = P_DEAD;                       // the anti-message itself is consumed
          v.outbox[(b.seq & S_IDX) - 1u].key = KEY_MAX;
          if (present != NIL) {                           // erase (queue.hpp:87-90)
            v.pflag[present] = P_DEAD;
            if (out_idx != NIL) v.outbox[out_idx].key = KEY_MAX;
            present = NIL; out_idx = NIL;
            ++s_annih;
          }
        } else if (present == NIL) {
          present = b.slot;                               // insert (queue.hpp:92)
          out_idx = (b.seq & S_IDX) - 1u;
        } else {
          v.pflag[b.slot] = P_DEAD;                       // duplicate key: second insert ignored
          v.outbox[(b.seq & S_IDX) - 1u].key = KEY_MAX;
          ++s_dup;
        }
      }
      eo[i] = make_uint2(present, out_idx);
    }
    __syncwarp();
    // --- B. state restore ---
    u32 st[M::kStateWords];
    load_state<M::kStateWords>(v.state + (size_t)lp * M::kStateWords, st);
    bool dirty = false;
    snap_pos = warp_min_u32(snap_pos);
    if (snap_pos != NIL) {
      const u32 si = (seg[snap_pos].seq & S_IDX) - 1u;
#pragma unroll
      for (int w = 0; w < M::kStateWords; ++w) st[w] = v.snap[(size_t)si * M::kStateWords + w];
      dirty = true;
    }
    // --- C. execute ---
    bool mutated = false;
    u32 maxt1 = 0, n_exec = 0;
    for (u32 i = lane; i < n; i += kWarp) {
      const uint2 s = eo[i];
      if (s.x == NIL) continue;
      const u32 t = key_time(seg[i].key);
      if (t >= tbound) { T.requeue(s.x); continue; }
      u32 out_idx = s.y;
      if (out_idx == NIL) { out_idx = T.extra_idx(); eo[i].y = out_idx; }      // a re-executed event has no inbox position
      if (T.try_process(s.x, out_idx, st)) mutated = true;
      ++n_exec;
      maxt1 = t + 1u > maxt1 ? t + 1u : maxt1;
    }
    if (__any_sync(0xffffffffu, mutated)) {
      __syncwarp();
      if (lane == 0) {
        for (u32 i = 0; i < n; ++i) {
          const uint2 s = eo[i];
          if (s.x == NIL || key_time(seg[i].key) >= tbound) continue;
          T.process_seq(s.x, s.y, st, dirty);
        }
      }
    }
    s_proc += n_exec;
    maxt1 = warp_max_u32(maxt1);
    // --- finish: newest processed time, state, ticket, history validity ---
    if (lane == 0) {
      const u32 lt_old = v.lpw[lp].x;
      const u64 inv_old = v.inv ? v.inv[lp] : KEY_MAX;
      // recorded events behind the window that were valid so far: the rollback to m reaches them too (lazily)
      const bool behind = v.inv && m < inv_old && lt_old > tbound;
      const u64 low = m < inv_old ? m : inv_old;      // everything at or above `low` inside the window was undone
      u32 base = lt_old;
      if ((n_und || behind) && low != KEY_MAX) {
        const u32 lt1 = key_time(low) + 1u;          // what is still processed lies below `low`
        base = lt1 < lt_old ? lt1 : lt_old;
      }
      if (n_und) ++s_rb;
      if (v.inv && m < inv_old && (n_und || behind)) {      // the recorded history of this LP is invalid from m on (R13)
        if (inv_old == KEY_MAX) atomicAdd(&c->n_touched, 1u);
        v.inv[lp] = m;
      }
      v.lpw[lp] = make_uint2(maxt1 > base ? maxt1 : base, 0u);
      if (dirty) {
#pragma unroll
        for (int w = 0; w < M::kStateWords; ++w) v.state[(size_t)lp * M::kStateWords + w] = st[w];
      }
    }
  }
  s_proc = warp_sum_u32(s_proc); s_arr = warp_sum_u32(s_arr); s_annih = warp_sum_u32(s_annih); s_dup = warp_sum_u32(s_dup);
  if (lane == 0) {
    if (s_proc != s_arr) atomicAdd(&c->n_processed, (u64)s_proc - (u64)s_arr);   // k_plan counted every arrival once
    if (s_annih) atomicAdd(&c->n_annihilated, (u64)s_annih);
    if (s_dup) atomicAdd(&c->n_dups, (u64)s_dup);
    if (s_rb) atomicAdd(&c->n_rollbacks, (u64)s_rb);
  }
  if (T.err) atomicOr(&c->error, T.err);
  if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&c->n_fix_total, (u64)nf); atomicAdd(&c->n_fix_rounds, 1ull); }
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

#ifndef SSB_EMU
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
#else
__device__ __forceinline__ void mbar_init(u64*, u32) {}
__device__ __forceinline__ void mbar_expect_tx(u64*, u32) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(u64*, u32) {}
#endif

// Each block stages one tile of the source (a contiguous run of 32-byte records) in shared memory with ONE TMA bulk
// copy, builds the per-target histogram from the staged records, claims one range per target with a single global
// atomic, and scatters the records from shared memory with 256-bit stores. Tiles are sized so that every block of
// the persistent grid gets the same number of equally large tiles (no tail wave).
__global__ void __launch_bounds__(kRouteThreads) k_route(View v, int which, u32 peer, u32 filter) {
#ifdef SSB_EMU
  alignas(128) static thread_local unsigned char smem_route[1 << 20];      // (per host thread: tools/emu/multi.py runs one engine per thread)
#else
  extern __shared__ __align__(128) unsigned char smem_route[];
#endif
  __shared__ u64 s_bar;
  __shared__ u64 s_sent_min;
  __shared__ const Ev* s_seg[2 * MAX_WORLD + 2];     // the source = a concatenation of up to 2 * world record runs
  __shared__ u32 s_off[2 * MAX_WORLD + 3];
  __shared__ u32 s_nseg;
  Ctl* c = v.ctl;
  if (threadIdx.x == 0) {
    u32 ns = 0, tot = 0;
    auto add = [&](const Ev* p, u32 n) { if (n) { s_seg[ns] = p; s_off[ns] = tot; tot += n; ++ns; } };
    if (which == 0) add(v.anti_box, c->n_anti < v.cap_aux ? c->n_anti : v.cap_aux);
    else if (which == 1) { add(v.outbox, c->n_arr); add(v.outbox + v.cap_batch, c->n_extra < v.cap_aux ? c->n_extra : v.cap_aux); }
    else if (which == 2) {
      if (v.p2p) {
        // own mailbox, all sources, both kinds in one launch (the records carry their senders' sequence stamps: the
        // order of integration does not matter)
        const u32 par = c->xround & 1u;
        for (u32 kind = 0; kind < 2; ++kind) {
          const u32 cap = kind ? v.mb_cap_anti[v.rank] : v.mb_cap_send[v.rank];
          for (u32 r = 0; r < v.world; ++r) {
            if (r == v.rank) continue;
            u32 n = *mb_cnt(v, v.rank, par, r, kind);
            if (n > cap) { n = cap; if (blockIdx.x == 0) atomicOr(&c->error, E_SENDBUF); }
            add(mb_data(v, v.rank, par, r, kind), n);
          }
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
  // records [0, n_main) are the handler outputs of this round's arrivals: each raises its sender's newest processed
  // time to its send time + 1 = the receive time of the event that produced it + 1. (An event whose handler produced
  // nothing leaves no trace: it had no effect, so nothing ever needs to undo it. The LPs of irregular turns got their
  // value from k_fix_turn; their outputs' send times lie below it.)
  const u32 n_main = which == 1 ? c->n_arr : 0u;
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
  // sequence stamp of what this launch routes for the first time (records that were routed before keep theirs)
  const u32 my_stamp = (c->stamp_base + (which == 1 ? 1u : 0u)) << F_STAMP_SHIFT;
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
          if (t0 + i < n_main) atomicMax(&v.lpw[part_local(v, h.w)].x, s_tile[i].send + 1u);
          if (dst >= v.n_lp) { atomicOr(&c->error, E_RANGE); continue; }
          const u32 tm = key_time(key);
          const u32 owner = v.world > 1 ? part_owner(v, dst) : v.rank;
          u32 t;
          if (tm >= v.finish_time && which != 3) {
            t = rs;      // never executed nor committed (R8): dropped where it is produced, not shipped to its owner
          } else if (owner != v.rank) {
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
        // the region of the peer's mailbox (or send buffer) this partition fills has one writer: a local counter
        s_base[t] = atomicAdd(send_kind && v.p2p ? &c->send_anti[r] : &c->send_count[r], n);
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
      const u32 fl = (qb.w & F_CANCEL) | ((qb.w >> F_STAMP_SHIFT) ? (qb.w & ~((1u << F_STAMP_SHIFT) - 1u)) : my_stamp);
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
        if (which == 3 && v.replay) {
          // history load (--diff_repeat, R13): the event was processed in the recorded run; its recorded sent-log
          // entry becomes its meta record
          const uint4 sr = v.hist_sent[t0 + i];
          store_meta(v.meta + slot, Meta{sr.x, NIL, ((u64)sr.z << 32) | sr.y});
          v.pflag[slot] = P_PROCESSED | P_HISTORY;
          atomicMax(&v.lpw[part_local(v, qa.z)].x, tm + 1u);
        } else {
          v.pflag[slot] = 0;
        }
      } else {
        const u32 r = t - rs - 1;
        const u32 p = s_base[t] + rnk[it];
        if (p >= (v.p2p ? (send_kind ? v.mb_cap_anti[r] : v.mb_cap_send[r]) : v.cap_send)) { atomicOr(&c->error, E_SENDBUF); continue; }
        if (v.p2p) {
          // straight into the destination GPU's mailbox over NVLink
          st256(mb_data(v, r, c->xround & 1u, v.rank, send_kind) + p, qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, fl);
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
  if (blockIdx.x == 0 && threadIdx.x == 0) { c->load_count = n; c->stamp_base += 2u; }
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
  if (blockIdx.x == 0 && threadIdx.x == 0) { c->load_count = n; c->stamp_base += 2u; }
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
  if (blockIdx.x == 0 && threadIdx.x == 0) { v.ctl->load_count = n; v.ctl->stamp_base += 2u; }
}

// state-change what-if query: the LP's recorded history is invalid from `key` on
__global__ void k_invalidate_from(View v, u32 lp, u64 key) {
  if (blockIdx.x || threadIdx.x) return;
  if (key < v.inv[lp]) {
    if (v.inv[lp] == KEY_MAX) v.ctl->n_touched += 1u;
    v.inv[lp] = key;
  }
}

// after the history load: every recorded event counts as consumed (processed); the last bucket with history
__global__ void k_after_history(View v) {
  Ctl* c = v.ctl;
  for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < v.nb; b += gridDim.x * blockDim.x) {
    const u32 f = v.fill[b];
    v.consumed[b] = f;
    if (f) atomicMax(&c->hist_end, b + 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) c->load_count = 0;
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
__global__ void k_fill_u64(u64* p, u64 val, size_t n) {
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
