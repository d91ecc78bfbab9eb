// ssb_device.cuh — device-side data layout of the Time Warp core (sm_100a).
//
// Layout in HBM:
//   * event pool   : Ev[max_events], 32-byte records = one DRAM sector each, carved into chunks of
//                    2^ch_log slots; a calendar bucket (receive_time / bucket_ticks) owns a list of chunks
//                    (chunk_tab) and is append-only, so "pending set" == the unconsumed tail of the buckets and
//                    "processed events" == the consumed head. Event records are immutable once routed.
//   * per-slot meta: Meta[max_events] (16 B: the sent-log entry — what set_cancel keeps in the reference,
//                    queue.hpp:151-157 — and the snapshot index) + pflag[max_events] (1 B: processed / dead / history).
//   * per-LP       : lpw {time of the newest processed event + 1, irregular-turn ticket of this round}, state words,
//                    inv (exact-differential repeat only: key from which the recorded history of the LP is invalid).
//   * per-round    : a hash set of the window's (LP, key) pairs (duplicate detection), outboxes, and — only for the
//                    LPs whose turn is not plain — sorted (key, slot, arrival) segments. Event records are never
//                    copied or sorted.
// There are no per-LP event lists or chains: a rollback finds the processed events of an LP by a coalesced scan over
// the consumed part of the window's buckets (k_fix_count / k_fix_scatter).
#pragma once
#include <cstdint>
#ifdef SSB_EMU
#include "cuda_emu.h"      // tools/emu: sequential host emulation of the kernels (developer tool, never shipped)
#else
#include <cuda_runtime.h>
#endif

namespace ssb {

typedef unsigned long long u64;
typedef unsigned int u32;

#ifdef SSB_EMU
constexpr u32 kWarp = 1;          // the emulation runs every kernel as one block of one thread
#else
constexpr u32 kWarp = 32;
#endif

constexpr u32 NIL = 0xFFFFFFFFu;
constexpr u64 KEY_MAX = ~0ull;
constexpr u32 F_CANCEL = 1u;      // anti-message (sim_event_base::is_cancel_, sim_obj.hpp:30)
// Ev::flags bits 4..31 = sequence stamp, given by the partition that first routes the record (0 = not routed yet): the
// order in which its sender emitted things. The merge applies the inbox items of one key in stamp order — the FIFO
// property rule R4 relies on ("a sender's positive always precedes its anti-message") — so it does not matter in
// which order records are appended to a bucket or cross between GPUs.
constexpr u32 F_STAMP_SHIFT = 4;
constexpr unsigned char P_PROCESSED = 1;   // pflag: handler has run and the result stands
constexpr unsigned char P_DEAD = 2;        // pflag: annihilated, duplicate, or a consumed anti-message
constexpr unsigned char P_HISTORY = 4;     // pflag: loaded from the exact-differential history store and not re-executed since
constexpr u32 TAG_START = 1u, TAG_END = 2u;
constexpr int MAX_WORLD = 16;
constexpr int RQ_LEN = 64;

// error bits (Ctl::error)
constexpr u32 E_POOL = 1u;        // event pool exhausted
constexpr u32 E_RANGE = 2u;       // time / id does not fit the packed record
constexpr u32 E_BATCH = 4u;       // aux outbox / fix segment overflow
constexpr u32 E_SNAP = 8u;        // snapshot ring overflow
constexpr u32 E_HORIZON = 16u;    // bucket index beyond the calendar
constexpr u32 E_SENDBUF = 32u;    // remote send buffer overflow
constexpr u32 E_INTERNAL = 64u;   // invariant violated
constexpr u32 E_COMM = 256u;      // a peer did not reach the GVT exchange in time
constexpr u32 E_STALL = 128u;     // GVT did not advance for kStallRounds rounds
constexpr u32 kStallRounds = 100000u;

// Event record: timestamp order of the reference (util/timestamp.hpp:53-59) is the u64 order of `key`.
struct __align__(32) Ev {
  u64 key;     // receive_time << 32 | id
  u32 dst;     // destination LP
  u32 src;     // source LP, NIL = -1
  u32 send;    // send_time
  u32 p0, p1;  // model payload
  u32 flags;
};
static_assert(sizeof(Ev) == 32, "Ev must be one 32-byte sector");

struct __align__(16) Meta {
  u32 sent_dst;  // sent-log: destination LP of the event this one produced, NIL = none
  u32 snap;      // index into the snapshot ring of the pre-state, NIL = handler left the state unchanged
  u64 sent_key;  // sent-log: key of the produced event
};
static_assert(sizeof(Meta) == 16, "Meta is one 16-byte access");

// entry of an irregular LP's segment (k_fix_*): sorted by (key, order of arrival)
constexpr u32 S_ANTI = 0x80000000u;     // the arrival is an anti-message
constexpr u32 S_UNDONE = 0x40000000u;   // a processed event that the rollback undid (it was in the queue before every arrival)
constexpr u32 S_IDX = 0x3FFFFFFFu;      // arrival: arrival index + 1; undone: snapshot index + 1 (0 = no snapshot)
struct __align__(16) Sorted {
  u64 key;
  u32 slot;      // calendar slot of the event
  u32 seq;       // S_ANTI | S_UNDONE | S_IDX
};
__device__ __forceinline__ u32 sorted_ord(u32 seq) { return (seq & S_UNDONE) ? 0u : (seq & S_IDX); }

struct Ctl {
  u64 local_min;      // min key over this partition's unconsumed events
  u64 gvt;            // global virtual time (packed key)
  u64 beyond_min;     // min key of events dropped because receive_time >= finish_time
  u32 done, error;
  u32 cur_abs, bound_abs, commit_next;
  u32 hi_bound;       // largest window bound so far: no event beyond bucket hi_bound has been processed by this run
  u32 hist_end;       // exact-differential repeat: buckets [0, hist_end) hold recorded history
  u32 n_touched;      // ... number of LPs whose history has been invalidated from some key on (0 = nothing to re-execute)
  u32 n_arr, n_anti, n_extra, commit_blocks, exec_blocks;
  u32 n_fix, n_ent, n_scan, scan_total;
  u32 hmask_round;    // size - 1 of the hash-set region this round uses
  u32 fix_min_t;      // smallest receive time among the arrivals that flagged an LP this round: no rollback of the round
                      // reaches below it, so the scans of the consumed bucket heads start at its bucket
  u32 n_gather, n_commit, commit_total;
  u32 free_top, free_low;
  u32 snap_head, snap_tail;
  u32 rq_head, rq_tail;
  u32 rq_bound[RQ_LEN], rq_snap[RQ_LEN];
  u32 round, load_count;
  u32 stamp_base;     // sequence stamps: +2 per round and per load; k_route(anti) stamps with it, k_route(main) with +1
  u32 xround;         // exchange rounds (global synchronisations) so far: parity of the mailbox halves and GVT slots
  u32 send_count[MAX_WORLD];      // events sent to each partition since the last global synchronisation
  u32 send_anti[MAX_WORLD];       // ... anti-messages (peer-memory exchange: they travel in their own mailbox region)
  u32 recv_count[MAX_WORLD];
  u64 prev_gvt;       // stall detection of the graph's loop
  u32 stall;
  u32 where;          // first internal invariant that failed (1 commit of an unprocessed event, 4 chunk never
                      // published, 5 arrival below GVT, 6 fix segment longer than counted)
  u32 where_arg;      // its argument (time / LP / bucket)
  u32 epoch;          // number of resets: keeps the GVT-exchange stamps of consecutive simulations apart
  u64 sent_min;       // peer-memory exchange: smallest bucket-start key of the events written to peers this round
  // statistics
  u64 n_processed, n_rollbacks, n_undone, n_antis, n_annihilated, n_dups, n_committed, n_beyond, n_remote;
  u64 n_fix_total;    // irregular LP turns so far
  u64 n_fix_rounds;   // rounds with at least one irregular turn
  u64 log_count, log_dropped;
  u64 log_stable;     // log_count at the start of the current round: everything below is completely written
  u64 digest[4];
};

struct View {
  // configuration
  u32 rank, world;
  u32 part_mode;          // 0: owner = lp % world, local = lp / world; 1: tables
  const u32* part_owner;  // [n_lp]
  const u32* part_local;  // [n_lp]
  u32 n_lp, n_lp_local;
  u32 nb;                 // calendar buckets (absolute index, no wrap: nb*bucket_ticks >= finish_time)
  u32 ch_log, maxc;       // chunk size log2, chunk-table row length
  u32 route_slots;        // buckets covered by k_route's shared-memory table
  u32 bucket_ticks, finish_time, window;
  u32 cap_batch, cap_aux, cap_snap, cap_send, cap_fix;
  u64 cap_log;
  u32 state_words;
  u32 use_graph;          // 1 = the kernels run inside the window-loop graph and drive its conditional nodes
  u32 sub;                // window index since the last global synchronisation (0 = the first; see ssb_set_sync_interval)
  u32 sync;               // windows per global synchronisation (1 = every window is synchronised)
  u32 hash_hi, hash_cap;  // hash-set region of a round: >= 2 x arrivals, hash_hi x while the region stays below hash_cap slots
  cudaGraphConditionalHandle h_loop, h_fix;
  // event pool
  Ev* ev;
  Meta* meta;
  unsigned char* pflag;   // [max_events] P_PROCESSED / P_DEAD / P_HISTORY
  u32* chunk_tab;   // [nb * maxc]
  u32* free_stack;  // [n_chunks]
  u32* fill;        // [nb] events appended
  u32* consumed;    // [nb] events already integrated
  // per LP
  uint2* lpw;       // {receive_time of the newest processed event + 1 (0 = none), ticket of an irregular turn in this
                    //  round: 0 = plain, 0xffffffff = being issued, else index into the fix table + 1}
  u32* state;       // [n_lp_local * state_words]
  u64* inv;         // exact-differential repeat only (else null): recorded events of the LP with key >= inv[lp] are invalid
  // per round
  u32* hset;           // [hcap] 32-bit fingerprints of the window's (LP, key) pairs; all zero between rounds
  u32 hmask;           // hcap - 1 (hcap is a power of two >= 2 * cap_batch)
  // irregular turns (anti-message, duplicate, straggler, state-mutating handler, invalidated history)
  u32* fixlp;          // [n_lp_local] fix table: LP of ticket f
  u32* fixcnt;         // [n_lp_local] entries counted for the LP (upper bound)
  u32* fixstart;       // [n_lp_local] start of its segment in `sorted`
  u32* fixpos;         // [n_lp_local] fill cursor of the segment
  u32* fixund;         // [n_lp_local] processed events undone by this round's rollback
  u64* fixmin;         // [n_lp_local] smallest arrival key = the key the LP rolls back to
  Sorted* sorted;      // [cap_fix] segments of the irregular LPs
  uint2* ent_out;      // [cap_fix] per segment entry: {slot of the event that survives the merge under this key, its outbox index}
  Ev* outbox;          // [cap_batch + cap_aux]: entry i = output of arrival i, then re-execution outputs
  Ev* anti_box;        // [cap_aux] anti-messages of this round
  uint4* gather;       // [window] {bucket, start, count, out_off}
  uint4* scan;         // [nb] consumed parts of the buckets a rollback can reach {bucket, 0, count, off}
  uint4* commit;       // [nb] {bucket, count, off, 0}
  Ev* sendbuf;         // [world * cap_send]
  Ev* recvbuf;         // [world * cap_send]
  // peer-memory exchange (NVLink P2P): every partition owns a mailbox block, mapped into all peers with CUDA IPC:
  //   header   256 bytes: {cap_send, cap_anti, GVT slot offset, record offset} of THIS mailbox — partitions may be
  //            configured with different capacities (e.g. a --diff_repeat run sizes them by the rank's history)
  //   counts   [2 parities][world sources][2 kinds]  one 128-byte line each (kind 0 = positives, 1 = anti-messages)
  //   GVT slot [2 parities][world sources]           u64 = round << 32 | local minimum time
  //   records  [2 parities][world sources] x { cap_send positives, cap_anti anti-messages }
  // A sender appends straight into its region of the destination's mailbox (a region has ONE writer, so positions
  // are claimed with a LOCAL atomic on the sender's own counter; plain stores over NVLink). At the start of the next
  // round every partition publishes its counts, writes its GVT candidate into all peers' slots and waits for theirs
  // (k_gvt): that exchange is the GVT min-reduction AND the barrier that makes the appended events visible.
  u32 p2p;                       // 0 = NCCL send/recv exchange, 1 = peer-memory exchange
  u32 cap_anti;                  // anti-message records per (parity, source) of the own mailbox
  u32 mb_cap_send[MAX_WORLD], mb_cap_anti[MAX_WORLD];     // geometry of every partition's mailbox (from its header)
  u32 mb_gvt_off[MAX_WORLD], mb_data_off[MAX_WORLD];      // byte offsets inside the block
  char* peer_base[MAX_WORLD];    // mailbox blocks of all partitions (own block at [rank])
  Ev* log;             // [cap_log] committed records
  uint4* log_sent;     // [cap_log] history store (--diff_init): sent-log entry of each committed event {dst, key lo, key hi, 0}
  uint4* hist_sent;    // [cap_batch] staging of the sent-log entries while history is loaded (--diff_repeat)
  u32 replay;          // 1 = a recorded history is being loaded: routed events become processed history events
  u32* snap;           // [cap_snap * state_words]
  Ctl* ctl;
};

__device__ __forceinline__ void internal_error(const View& v, u32 where, u32 arg) {
  if (atomicCAS(&v.ctl->where, 0u, where) == 0u) v.ctl->where_arg = arg;
  atomicOr(&v.ctl->error, E_INTERNAL);
}
__device__ __forceinline__ u32 key_time(u64 k) { return (u32)(k >> 32); }
__device__ __forceinline__ u32 key_id(u64 k) { return (u32)k; }
__device__ __forceinline__ u64 make_key(u32 t, u32 id) { return ((u64)t << 32) | id; }

__device__ __forceinline__ u32 part_owner(const View& v, u32 lp) {
  return v.part_mode == 0 ? lp % v.world : v.part_owner[lp];
}
__device__ __forceinline__ u32 part_local(const View& v, u32 lp) {
  return v.part_mode == 0 ? lp / v.world : v.part_local[lp];
}

// 32-byte records move as ONE 256-bit access per thread (LDG/STG.256 on sm_100a): a warp touches every DRAM sector
// once, instead of twice with two 128-bit halves.
#ifndef SSB_EMU
__device__ __forceinline__ void ld256(const void* p, u32 (&x)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
               : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(void* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
// streaming variants (L2 evict-first): for data that is not touched again soon — committed buckets, the log, meta
// records, events appended to future buckets — so that they do not push the current window out of L2
__device__ __forceinline__ void ld256_ef(const void* p, u32 (&x)[8]) {
  asm volatile("ld.global.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
               : "l"(p) : "memory");
}
__device__ __forceinline__ void st256_ef(void* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) {
  asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void st128_ef(void* p, u32 a, u32 b, u32 c, u32 d) {
  // (the L2::evict_first qualifier exists for 256-bit stores only; a 128-bit store takes the streaming hint)
  asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
#else
__device__ __forceinline__ void ld256(const void* p, u32 (&x)[8]) { for (int i = 0; i < 8; ++i) x[i] = static_cast<const u32*>(p)[i]; }
__device__ __forceinline__ void st256(void* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) {
  u32* q = static_cast<u32*>(p); q[0] = a; q[1] = b; q[2] = c; q[3] = d; q[4] = e; q[5] = f; q[6] = g; q[7] = h;
}
__device__ __forceinline__ void ld256_ef(const void* p, u32 (&x)[8]) { ld256(p, x); }
__device__ __forceinline__ void st256_ef(void* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) { st256(p, a, b, c, d, e, f, g, h); }
__device__ __forceinline__ void st128_ef(void* p, u32 a, u32 b, u32 c, u32 d) { u32* q = static_cast<u32*>(p); q[0] = a; q[1] = b; q[2] = c; q[3] = d; }
#endif
__device__ __forceinline__ Ev load_ev_ef(const Ev* p) {
  u32 x[8];
  ld256_ef(p, x);
  Ev e;
  e.key = ((u64)x[1] << 32) | x[0]; e.dst = x[2]; e.src = x[3];
  e.send = x[4]; e.p0 = x[5]; e.p1 = x[6]; e.flags = x[7];
  return e;
}
__device__ __forceinline__ void store_ev_ef(Ev* p, const Ev& e) {
  st256_ef(p, (u32)e.key, (u32)(e.key >> 32), e.dst, e.src, e.send, e.p0, e.p1, e.flags);
}
__device__ __forceinline__ Ev load_ev(const Ev* p) {
  u32 x[8];
  ld256(p, x);
  Ev e;
  e.key = ((u64)x[1] << 32) | x[0]; e.dst = x[2]; e.src = x[3];
  e.send = x[4]; e.p0 = x[5]; e.p1 = x[6]; e.flags = x[7];
  return e;
}
__device__ __forceinline__ void store_ev(Ev* p, const Ev& e) {
  st256(p, (u32)e.key, (u32)(e.key >> 32), e.dst, e.src, e.send, e.p0, e.p1, e.flags);
}
__device__ __forceinline__ Meta load_meta(const Meta* p) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  Meta m;
  m.sent_dst = q.x; m.snap = q.y; m.sent_key = ((u64)q.w << 32) | q.z;
  return m;
}
__device__ __forceinline__ void store_meta(Meta* p, const Meta& m) {
  *reinterpret_cast<uint4*>(p) = make_uint4(m.sent_dst, m.snap, (u32)m.sent_key, (u32)(m.sent_key >> 32));
}
// a meta record is not read again unless its LP rolls back (or the history store keeps it): evict-first
__device__ __forceinline__ void store_meta_ef(Meta* p, const Meta& m) {
  st128_ef(p, m.sent_dst, m.snap, (u32)m.sent_key, (u32)(m.sent_key >> 32));
}

__device__ __forceinline__ u32* mb_cnt(const View& v, u32 owner, u32 par, u32 src, u32 kind) {
  return reinterpret_cast<u32*>(v.peer_base[owner] + 256) + ((par * v.world + src) * 2u + kind) * 32u;
}
__device__ __forceinline__ u64* mb_gvt(const View& v, u32 owner, u32 par, u32 src) {
  return reinterpret_cast<u64*>(v.peer_base[owner] + v.mb_gvt_off[owner]) + par * v.world + src;
}
__device__ __forceinline__ Ev* mb_data(const View& v, u32 owner, u32 par, u32 src, u32 kind) {
  return reinterpret_cast<Ev*>(v.peer_base[owner] + v.mb_data_off[owner]) +
         (size_t)(par * v.world + src) * (v.mb_cap_send[owner] + v.mb_cap_anti[owner]) + (kind ? v.mb_cap_send[owner] : 0u);
}

// slot of position p in bucket b; the chunk must already exist
// (after an event-pool overflow — E_POOL — a bucket's fill count can point past its chunk list or at a chunk that was
// never allocated: such a position resolves into the sacrificial chunk 0, never outside the pool)
__device__ __forceinline__ u32 slot_of(const View& v, u32 b, u32 p) {
  const u32 j = p >> v.ch_log;
  u32 ch = j < v.maxc ? v.chunk_tab[(size_t)b * v.maxc + j] : 0u;
  if (ch == NIL) ch = 0u;
  return (ch << v.ch_log) | (p & ((1u << v.ch_log) - 1u));
}

__device__ __forceinline__ u64 mix64(u64 z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}
// hash of one committed record; identical to oracle/tw_oracle.cc record_hash (values widened as int64)
__device__ __forceinline__ u64 record_hash(const Ev& e, u32 tag) {
  u64 h = 0x9E3779B97F4A7C15ULL;
  h = mix64(h ^ (u64)key_id(e.key));
  h = mix64(h ^ (e.src == NIL ? ~0ull : (u64)e.src));
  h = mix64(h ^ (u64)e.send);
  h = mix64(h ^ (u64)e.dst);
  h = mix64(h ^ (u64)key_time(e.key));
  h = mix64(h ^ (u64)tag);
  return h;
}

}  // namespace ssb
