// ssb_device.cuh — device-side data layout of the Time Warp core (sm_100a).
//
// Layout in HBM:
//   * event pool   : Ev[max_events], 32-byte records = one DRAM sector each, carved into chunks of
//                    2^ch_log slots; a calendar bucket (receive_time / bucket_ticks) owns a list of chunks
//                    (chunk_tab) and is append-only, so "pending set" == the unconsumed tail of the buckets.
//                    Event records are immutable once routed.
//   * per-slot meta: Meta[max_events] (32 B: processed-chain link, sent-log entry — what set_cancel keeps in the
//                    reference, queue.hpp:151-157 — and snapshot index) + pflag[max_events] (1 B: processed / dead).
//   * per-LP       : cnt / lpw {segment start, length, chain head slot, its time} / lpflag / last_slot / last_time /
//                    state words.
//   * per-round    : arrival staging, per-LP sorted segments, outboxes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssb {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u32 NIL = 0xFFFFFFFFu;
constexpr u64 KEY_MAX = ~0ull;
constexpr u32 F_CANCEL = 1u;      // anti-message (sim_event_base::is_cancel_, sim_obj.hpp:30)
constexpr unsigned char P_PROCESSED = 1;   // pflag: handler has run and the result stands
constexpr unsigned char P_DEAD = 2;        // pflag: annihilated, duplicate, or a consumed anti-message
constexpr u32 TAG_START = 1u, TAG_END = 2u;
constexpr int MAX_WORLD = 16;
constexpr int RQ_LEN = 64;

// error bits (Ctl::error)
constexpr u32 E_POOL = 1u;        // event pool exhausted
constexpr u32 E_RANGE = 2u;       // time / id does not fit the packed record
constexpr u32 E_BATCH = 4u;       // aux outbox overflow
constexpr u32 E_SNAP = 8u;        // snapshot ring overflow
constexpr u32 E_HORIZON = 16u;    // bucket index beyond the calendar
constexpr u32 E_SENDBUF = 32u;    // remote send buffer overflow
constexpr u32 E_INTERNAL = 64u;   // invariant violated

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

struct __align__(32) Meta {
  u32 link;      // previous processed event of this LP (slot) or NIL
  u32 prevtime;  // receive_time of that previous event (so a stale link is never dereferenced)
  u32 sent_dst;  // sent-log: destination LP of the event this one produced, NIL = none
  u32 snap;      // index into the snapshot ring of the pre-state, NIL = handler left the state unchanged
  u64 sent_key;  // sent-log: key of the produced event
  u32 pad0, pad1;
};
static_assert(sizeof(Meta) == 32, "Meta must be one 32-byte sector");

struct __align__(16) Sorted {
  u64 key;
  u32 slot;
  u32 seq;       // arrival order (bits 0..30), bit 31 = anti-message
};

struct Ctl {
  u64 local_min;      // min key over this partition's unconsumed events
  u64 gvt;            // global virtual time (packed key)
  u64 beyond_min;     // min key of events dropped because receive_time >= finish_time
  u32 done, error;
  u32 cur_abs, bound_abs, commit_next;
  u32 n_arr, seg_cursor, n_anti, n_extra, n_heavy, n_fix, n_fixheavy, commit_blocks;
  u32 n_gather, n_commit, commit_total;
  u32 free_top, free_low;
  u32 snap_head, snap_tail;
  u32 rq_head, rq_tail;
  u32 rq_bound[RQ_LEN], rq_snap[RQ_LEN];
  u32 round, load_count;
  u32 send_count[MAX_WORLD];
  u32 recv_count[MAX_WORLD];
  u64 sent_min;       // peer-memory exchange: smallest bucket-start key of the events written to peers this round
  // statistics
  u64 n_processed, n_rollbacks, n_undone, n_antis, n_annihilated, n_dups, n_committed, n_beyond, n_remote;
  u64 log_count, log_dropped;
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
  u32 cap_batch, cap_aux, cap_snap, cap_send;
  u64 cap_log;
  u32 state_words;
  // event pool
  Ev* ev;
  Meta* meta;
  unsigned char* pflag;   // [max_events] P_PROCESSED / P_DEAD
  u32* chunk_tab;   // [nb * maxc]
  u32* free_stack;  // [n_chunks]
  u32* fill;        // [nb] events appended
  u32* consumed;    // [nb] events already integrated
  // per LP
  u32* cnt;
  uint4* lpw;       // this round: {segment start in `sorted`, segment length, chain head slot, its receive_time}
  u32* lpflag;      // this round: PL_* bits
  u32* last_slot;
  u32* last_time;
  u32* state;       // [n_lp_local * state_words]
  // per round
  Ev* tmp_work;        // [cap_batch] work records (event, flags = arrival index | anti << 31), arrival order
  u32* tmp_lp;         // [cap_batch]
  u32* tmp_rank;       // [cap_batch]
  Ev* work;            // [cap_batch] work records grouped by LP (segments)
  Sorted* sorted;      // [cap_batch] (key, slot, arrival) of a segment, sorted — only for long or irregular segments
  uint4* fix;          // [cap_batch] LPs that need the sequential turn {lp, seg_start, n, 0}
  uint4* fixheavy;     // [cap_batch / 16] same, segments that still need a warp-level sort
  uint4* heavy;        // [cap_batch / 64] LPs with long segments {lp, seg_start, n, 0}
  Ev* outbox;          // [cap_batch + cap_aux]: main region aligned with `sorted`, then re-execution outputs
  Ev* anti_box;        // [cap_aux] anti-messages of this round
  uint4* gather;       // [window] {bucket, start, count, out_off}
  uint4* commit;       // [nb] {bucket, count, off, 0}
  Ev* sendbuf;         // [world * cap_send]
  Ev* recvbuf;         // [world * cap_send]
  // peer-memory exchange (NVLink P2P): every partition owns a mailbox = [2 parities][world sources] x {count, records}.
  // A sender appends straight into its region of the destination's mailbox (system-scope atomic on the count, then
  // plain stores over NVLink); the GVT all-reduce of the next round is the barrier that makes the appends visible.
  u32 p2p;                       // 0 = NCCL send/recv exchange, 1 = peer-memory exchange
  Ev* mbox_data;                 // own mailbox records [2][world][cap_send]
  u32* mbox_cnt;                 // own mailbox counts  [2][world] (one 128-byte line each)
  Ev* peer_data[MAX_WORLD];      // peers' mailboxes (mapped with CUDA IPC)
  u32* peer_cnt[MAX_WORLD];
  Ev* log;             // [cap_log] committed records
  u32* snap;           // [cap_snap * state_words]
  Ctl* ctl;
};

__device__ __forceinline__ u32 key_time(u64 k) { return (u32)(k >> 32); }
__device__ __forceinline__ u32 key_id(u64 k) { return (u32)k; }
__device__ __forceinline__ u64 make_key(u32 t, u32 id) { return ((u64)t << 32) | id; }

__device__ __forceinline__ u32 part_owner(const View& v, u32 lp) {
  return v.part_mode == 0 ? lp % v.world : v.part_owner[lp];
}
__device__ __forceinline__ u32 part_local(const View& v, u32 lp) {
  return v.part_mode == 0 ? lp / v.world : v.part_local[lp];
}

__device__ __forceinline__ Ev load_ev(const Ev* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Ev e;
  e.key = ((u64)a.y << 32) | a.x; e.dst = a.z; e.src = a.w;
  e.send = b.x; e.p0 = b.y; e.p1 = b.z; e.flags = b.w;
  return e;
}
__device__ __forceinline__ Meta load_meta(const Meta* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Meta m;
  m.link = a.x; m.prevtime = a.y; m.sent_dst = a.z; m.snap = a.w;
  m.sent_key = ((u64)b.y << 32) | b.x; m.pad0 = 0; m.pad1 = 0;
  return m;
}
__device__ __forceinline__ void store_meta(Meta* p, const Meta& m) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(m.link, m.prevtime, m.sent_dst, m.snap);
  q[1] = make_uint4((u32)m.sent_key, (u32)(m.sent_key >> 32), 0u, 0u);
}
__device__ __forceinline__ void store_ev(Ev* p, const Ev& e) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4((u32)e.key, (u32)(e.key >> 32), e.dst, e.src);
  q[1] = make_uint4(e.send, e.p0, e.p1, e.flags);
}

// slot of position p in bucket b; the chunk must already exist
__device__ __forceinline__ u32 slot_of(const View& v, u32 b, u32 p) {
  u32 ch = v.chunk_tab[(size_t)b * v.maxc + (p >> v.ch_log)];
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
