// ssb_models.cuh — device twins of the application event handlers.
//
// A model provides
//   kStateWords              size of the LP state in 32-bit words
//   kUsesState               false: the handler neither reads nor writes the state (k_exec then skips its gather)
//   kExecMinBlocks           __launch_bounds__ hint of k_exec for this handler (resident blocks per SM to aim for)
//   struct Data              read-only model data (device pointers)
//   early(d, e)              the ONE word of model data the handler gathers that depends on the event alone (PHOLD: the
//                            table entry; TrafficSim: the next route node). k_exec issues it together with the other
//                            gathers of the event, before anything else is known, so that the latencies overlap.
//   handle(d, e, state, early, out) App::event_handler: returns true and fills `out` when one new event is
//                            produced; may modify `state` (kept only when an event is produced, which
//                            is what runner.hpp:552-558 does — quirk Q2 of SURVEY.md)
//   tag(e)                   START / END suffix of the committed line (sim_obj.hpp:73-75)
// Both shipped ScaleSim applications produce at most one event per handled event.
#pragma once
#include "ssb_device.cuh"

namespace ssb {

// ---------------------------------------------------------------------------------------------
// PHOLD — src/phold/phold.hpp:256-281
//   hops' = hops + 1; idx = (int)(id + hops') % RAND_TABLE_SIZE
//   recv' = recv + LATENCY[idx] + LOOK_AHEAD; dst' = REMOTE[idx] == 1 ? LATENCY[idx] % NUM_LP : dst
//   src' = dst; send' = recv; id' = id.   State {long id} is returned unchanged.
// The two host tables are packed into one u32 table: latency << 1 | remote (latency < 2^31 is
// checked when the tables are uploaded).
// ---------------------------------------------------------------------------------------------
struct PholdModel {
  static constexpr int kStateWords = 2;
  static constexpr bool kUsesState = false;      // phold::event_handler returns the state it was given (phold.hpp:279-280)
  static constexpr int kExecMinBlocks = 8;       // k_exec fits 32 registers with this handler: 8 resident blocks per SM
  struct Data {
    const u32* table;
    u32 table_size;
    u32 lookahead;
    u32 num_lp;
  };
  __device__ static __forceinline__ u32 early(const Data& d, const Ev& e) {
    // (int)(ev_id + (long)num_hops) % RAND_TABLE_SIZE — id + hops < 2^31 is asserted at load (quirk Q8)
    return __ldg(d.table + (key_id(e.key) + e.p0 + 1u) % d.table_size);
  }
  __device__ static __forceinline__ bool handle(const Data& d, const Ev& e, u32* state, u32 t, Ev& out, u32& err) {
    (void)state;
    u32 hops = e.p0 + 1u;
    u32 lat = t >> 1;
    u64 recv = (u64)key_time(e.key) + lat + d.lookahead;
    if (recv >= 0x80000000ull) { err |= E_RANGE; recv = 0x7FFFFFFFull; }
    out.key = make_key((u32)recv, key_id(e.key));
    out.dst = (t & 1u) ? lat % d.num_lp : e.dst;
    out.src = e.dst;
    out.send = key_time(e.key);
    out.p0 = hops;
    out.p1 = 0;
    out.flags = 0;
    return true;
  }
  __device__ static __forceinline__ u32 tag(const Data&, const Ev& e) {
    return e.src == NIL ? TAG_START : TAG_END;   // phold::Event::end() is always true (phold.hpp:89)
  }
};

// ---------------------------------------------------------------------------------------------
// PHOLD with mutable LP state — a test model (no reference application behaves like this; the
// parity target is the oracle's sequential execution). The LP state {count, chain} changes with
// every event and feeds back into the produced event, so any ordering or rollback mistake
// changes the committed records.
//   count' = count + 1; chain' = mix(chain ^ id ^ recv)
//   as PHOLD, except idx = (id + hops' + (chain' & 7)) % RAND_TABLE_SIZE and p1' = count'
// ---------------------------------------------------------------------------------------------
struct PholdStatefulModel {
  static constexpr int kStateWords = 2;
  static constexpr bool kUsesState = true;
  static constexpr int kExecMinBlocks = 1;
  typedef PholdModel::Data Data;
  __device__ static __forceinline__ u32 early(const Data&, const Ev&) { return 0u; }      // its gather depends on the state
  __device__ static __forceinline__ bool handle(const Data& d, const Ev& e, u32* state, u32, Ev& out, u32& err) {
    u32 count = state[0] + 1u;
    u32 chain = (u32)mix64(((u64)state[1] << 32) ^ e.key);
    state[0] = count;
    state[1] = chain;
    u32 hops = e.p0 + 1u;
    u32 idx = (key_id(e.key) + hops + (chain & 7u)) % d.table_size;
    u32 t = __ldg(d.table + idx);
    u32 lat = t >> 1;
    u64 recv = (u64)key_time(e.key) + lat + d.lookahead;
    if (recv >= 0x80000000ull) { err |= E_RANGE; recv = 0x7FFFFFFFull; }
    out.key = make_key((u32)recv, key_id(e.key));
    out.dst = (t & 1u) ? lat % d.num_lp : e.dst;
    out.src = e.dst;
    out.send = key_time(e.key);
    out.p0 = hops;
    out.p1 = count;
    out.flags = 0;
    return true;
  }
  __device__ static __forceinline__ u32 tag(const Data&, const Ev& e) { return e.src == NIL ? TAG_START : TAG_END; }
};

// ---------------------------------------------------------------------------------------------
// TrafficSim — src/trafficsim/traffic_sim.hpp:279-342
// Event payload: p0 = offset of destination() in the route pool, p1 = route nodes left (>= 1).
// LP state: 8 words = one 32-byte sector: [0] = id, [1] = number of out-roads n, then
//   n <= 2 : the roads inline, n x {dst, speed km/h, length m}            (grids: one gather per event)
//   n >  2 : [2] = index of the LP's first road in the road table (model data slot 1: {dst, speed, length} triples,
//            the LP's roads contiguous, in the order of the state's vectors) — any junction degree
//   no event when the route is exhausted (p1 == 1) or no road leads to the next node (:286-306)
//   arrival' = len*3600/speed/1000 + arrival + 5   (speed != 0, left-to-right truncating division, :314-316)
//            = len*3600 + arrival                   (speed == 0, :318-319)
//   departure' = arrival; source' = this node; the state is returned unchanged.
// ---------------------------------------------------------------------------------------------
struct TrafficModel {
  static constexpr int kInlineRoads = 2;
  static constexpr int kStateWords = 2 + 3 * kInlineRoads;
  static constexpr bool kUsesState = true;
  static constexpr int kExecMinBlocks = 1;
  struct Data {
    const u32* route;   // route pool
    u32 route_len;
    const u32* roads;   // road table of the junctions with more than kInlineRoads out-roads
    u32 roads_len;      // in words
  };
  __device__ static __forceinline__ u32 early(const Data& d, const Ev& e) {
    return (e.p1 > 1u && e.p0 + 1u < d.route_len) ? __ldg(d.route + e.p0 + 1u) : 0u;      // the next route node
  }
  __device__ static __forceinline__ bool handle(const Data& d, const Ev& e, u32* state, u32 next, Ev& out, u32& err) {
    if (e.p1 <= 1u) return false;
    if (e.p0 + 1u >= d.route_len) { err |= E_INTERNAL; return false; }
    const u32 n = state[1];
    // first road that leads to `next`, as std::find (traffic_sim.hpp:299-306)
    bool found = false;
    u64 speed = 0, len = 0;
    if (n <= (u32)kInlineRoads) {
      // static indices only, so the state words stay in registers
#pragma unroll
      for (int r = 0; r < kInlineRoads; ++r) {
        if (!found && (u32)r < n && state[2 + 3 * r] == next) {
          found = true;
          speed = state[2 + 3 * r + 1];
          len = state[2 + 3 * r + 2];
        }
      }
    } else {
      const u32 first = state[2];
      if ((u64)3 * ((u64)first + n) > d.roads_len) { err |= E_INTERNAL; return false; }
      for (u32 r = 0; r < n && !found; ++r) {
        const u32* rd = d.roads + 3u * (first + r);
        if (__ldg(rd) == next) { found = true; speed = __ldg(rd + 1); len = __ldg(rd + 2); }
      }
    }
    if (!found) return false;
    u64 arrival = key_time(e.key);
    // len * 60 * 60 / speed / 1000, left to right, truncating (traffic_sim.hpp:314-316); 32-bit arithmetic when the
    // product fits (roads shorter than 1 193 km), identical result
    u64 recv;
    if (speed == 0) recv = len * 60ull * 60ull + arrival;
    else if (len < 1193046ull) recv = (u64)((u32)len * 3600u / (u32)speed / 1000u) + arrival + 5ull;
    else recv = len * 60ull * 60ull / speed / 1000ull + arrival + 5ull;
    if (recv >= 0x80000000ull) { err |= E_RANGE; recv = 0x7FFFFFFFull; }
    out.key = make_key((u32)recv, key_id(e.key));
    out.dst = next;
    out.src = e.dst;          // destinations_.front() == this LP
    out.send = (u32)arrival;
    out.p0 = e.p0 + 1u;
    out.p1 = e.p1 - 1u;
    out.flags = 0;
    return true;
  }
  __device__ static __forceinline__ u32 tag(const Data&, const Ev& e) {
    if (e.src == NIL) return TAG_START;
    return e.p1 == 1u ? TAG_END : 0u;   // end() == (destinations_.size() == 1), traffic_sim.hpp:77
  }
};

}  // namespace ssb
