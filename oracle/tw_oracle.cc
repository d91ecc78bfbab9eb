/*
 * tw_oracle.cc — CPU restatement of ScaleSim's Time Warp core and its two applications.
 *
 * TEST INFRASTRUCTURE ONLY (see tw_oracle.h). Plain C++17, no third-party dependencies.
 * Every block cites the reference file:line it restates (paths relative to the reference root).
 *
 * Parity status: PINNED. tests/test_oracle_kat.py replays test/medium/logical_process_test.cc,
 * tests/test_oracle_golden.py compares whole-simulation output with traffic/README.md:119-126 and
 * with outputs of the unmodified reference (oracle/_ref) stored under tests/golden/.
 */
#include "tw_oracle.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

/* ---- util/timestamp.hpp:17-83 : (time,id), lexicographic ---- */
struct TS {
  long time, id;
  bool operator<(const TS& r) const { return time < r.time || (time == r.time && id < r.id); }
  bool operator==(const TS& r) const { return time == r.time && id == r.id; }
  static TS max() { return {std::numeric_limits<long>::max(), std::numeric_limits<long>::max()}; }
  static TS zero() { return {0, 0}; }
  static TS null() { return {-1, -1}; }
};

/* ---- event: union of phold::Event (phold.hpp:62-107) and traffic_sim::Event (traffic_sim.hpp:36-96) ---- */
struct Ev {
  long id = -1, src = -1, dst = -1, recv = -1, send = -1;
  int hops = 0;                                      /* phold */
  std::shared_ptr<const std::vector<long> > route;   /* traffic: whole route; remaining = route[rpos..] */
  long rpos = 0;
  bool cancel = false;
  bool end() const { return route ? (long)route->size() - rpos == 1 : true; } /* phold.hpp:89, traffic_sim.hpp:77 */
  TS key() const { return {recv, id}; }
};

inline two_record to_record(const Ev& e) {
  /* sim_obj.hpp:66-77 */
  two_record r;
  r.id = e.id; r.src = e.src; r.send = e.send; r.dst = e.dst; r.recv = e.recv;
  r.tag = (e.src == -1) ? 1 : (e.end() ? 2 : 0);
  return r;
}

inline uint64_t mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}
inline uint64_t record_hash(const two_record& r) {
  uint64_t h = 0x9E3779B97F4A7C15ULL;
  h = mix64(h ^ (uint64_t)r.id);
  h = mix64(h ^ (uint64_t)r.src);
  h = mix64(h ^ (uint64_t)r.send);
  h = mix64(h ^ (uint64_t)r.dst);
  h = mix64(h ^ (uint64_t)r.recv);
  h = mix64(h ^ (uint64_t)r.tag);
  return h;
}
inline void digest_add(uint64_t d[4], const two_record& r) {
  uint64_t h = record_hash(r);
  d[0] += 1; d[1] += h; d[2] ^= h; d[3] += mix64(h + 0xD1B54A32D192ED03ULL);
}

/* ---- eventq (queue.hpp:25-241) ---- */
struct EventQ {
  std::map<TS, Ev> map_, can_map_;
  std::vector<std::pair<TS, Ev> > buffer_;
  TS released_time_ = TS::zero(), released_cn_time_ = TS::zero(), outputted_time_ = TS::zero();

  void buffering(const Ev& ev) { buffer_.emplace_back(ev.key(), ev); }          /* queue.hpp:76-80 */

  TS merge_buffer(std::vector<Ev>& new_cancels) {                                /* queue.hpp:82-108 */
    TS min_stmp = TS::max();
    for (auto& it : buffer_) {
      if (it.second.cancel) {
        if (map_.erase(it.first)) min_stmp = std::min(min_stmp, it.first);
      } else {
        map_.insert(it);                       /* duplicate key: second insert ignored */
        min_stmp = std::min(min_stmp, it.first);
      }
    }
    buffer_.clear();
    auto can_it = can_map_.lower_bound(min_stmp);
    auto first = can_it;
    for (; can_it != can_map_.end(); ++can_it) {
      Ev c = can_it->second;
      c.cancel = true;
      new_cancels.push_back(c);
    }
    can_map_.erase(first, can_map_.end());
    return min_stmp;
  }
  void direct_insert(const Ev& ev) { map_.insert({ev.key(), ev}); }              /* queue.hpp:110-114 */

  bool increment(Ev& out, TS* local_time) {                                      /* queue.hpp:116-135 */
    if (*local_time == TS::max()) return false;
    auto it = map_.lower_bound(*local_time);
    if (it == map_.end()) { *local_time = TS::max(); return false; }
    out = it->second;
    if (++it != map_.end()) *local_time = it->first; else *local_time = TS::max();
    return true;
  }
  void set_cancel(const Ev& ev) {                                                /* queue.hpp:151-157 */
    Ev c = ev; c.cancel = true;
    can_map_.insert({TS{c.send, c.id}, c});
  }
  void release(const TS& to) {                                                   /* queue.hpp:159-167 */
    map_.erase(map_.lower_bound(released_time_), map_.lower_bound(to));
    released_time_ = to;
  }
  void release_cancel(const TS& to) {                                            /* queue.hpp:169-177 */
    can_map_.erase(can_map_.lower_bound(released_cn_time_), can_map_.lower_bound(to));
    released_cn_time_ = to;
  }
  template <class F> void std_out(const TS& to, F emit) {                        /* queue.hpp:203-211 */
    auto end = map_.lower_bound(to);
    for (auto it = map_.lower_bound(outputted_time_); it != end && it != map_.end(); ++it) emit(it->second);
    outputted_time_ = to;
  }
};

/* ---- stateq (queue.hpp:243-331); a state version is one integer ---- */
struct StateQ {
  std::map<TS, long> state_map;
  TS released_st_time = TS::zero();
  long get_state() const { auto it = state_map.end(); --it; return it->second; }  /* queue.hpp:273-278 */
  void push_back_state(long s, const TS& t) { state_map.insert({t, s}); }         /* queue.hpp:280-284 */
  void rollback_state(const TS& lb) { state_map.erase(state_map.lower_bound(lb), state_map.end()); } /* :286-290 */
  void release(const TS& to) {                                                    /* queue.hpp:292-302 (quirk Q1 kept) */
    state_map.erase(state_map.lower_bound(released_st_time), state_map.lower_bound(to));
    released_st_time = to;
  }
};

struct Scheduler;

/* ---- lp (logical_process.hpp:31-232) ---- */
struct LP {
  long id_ = 0;
  TS local_time_ = TS::max();
  EventQ eventq_;
  StateQ stateq_;
  long n_events = 0, n_cancels = 0;
  Scheduler* sched = nullptr;

  void buffer(const Ev& ev);                                                      /* :115-127 */
  void flush_buf(std::vector<Ev>& new_cancels) {                                  /* :129-157 (no diff_repeat) */
    local_time_ = std::min(local_time_, eventq_.merge_buffer(new_cancels));
    stateq_.rollback_state(local_time_);
    n_cancels += (long)new_cancels.size();
  }
  void direct_insert(const Ev& ev) {                                              /* :159-166 */
    eventq_.direct_insert(ev);
    local_time_ = std::min(local_time_, ev.key());
  }
  bool dequeue_event(Ev& out) {                                                   /* :168-173 */
    bool ok = eventq_.increment(out, &local_time_);
    if (ok) ++n_events;
    return ok;
  }
};

/* ---- scheduler (process_scheduler.hpp:26-100): lowest-timestamp-first ---- */
struct Scheduler {
  std::map<TS, long> ltsf_queue_;
  std::unordered_map<long, TS> index_;
  std::unordered_set<long> active_lp_;
  long dequeue() {                                                                /* :55-67 */
    auto it = ltsf_queue_.begin();
    if (it == ltsf_queue_.end() || it->first == TS::max()) return -1;
    long ret = it->second;
    index_.erase(ret);
    ltsf_queue_.erase(it);
    active_lp_.insert(ret);
    return ret;
  }
  void queue(const TS& t, long lp_id) {                                           /* :69-81 */
    auto f = index_.find(lp_id);
    if (f != index_.end()) {
      TS old_time = f->second;
      if (old_time < t) return;
      ltsf_queue_.erase(old_time);
      index_.erase(lp_id);
    }
    ltsf_queue_.insert({t, lp_id});
    index_.insert({lp_id, t});
  }
  TS min_locals() const {                                                         /* :83-90 */
    if (ltsf_queue_.empty()) return TS::max();
    return ltsf_queue_.begin()->first;
  }
};

void LP::buffer(const Ev& ev) {
  eventq_.buffering(ev);
  local_time_ = std::min(local_time_, ev.key());
  if (sched) sched->queue(local_time_, id_);
}

/* ---- PHOLD tables (phold.hpp:144-164) — libstdc++ default_random_engine = minstd_rand0 ---- */
const int kEffectiveDecimal = 100000;                   /* phold.hpp:54 */
const long kLookAhead = (long)(0.1 * kEffectiveDecimal);/* phold.hpp:55 */
const int kRandTableSize = 10000000;                    /* phold.hpp:56 */

void build_phold_tables(long seed, double lambda, double ratio, long n, int64_t* lat, int32_t* rem) {
  {
    int ex_seed = (int)seed;
    std::default_random_engine ex_generator(ex_seed);
    std::exponential_distribution<double> ex_distribution(lambda);
    for (long i = 0; i < n; ++i) lat[i] = (long)(ex_distribution(ex_generator) * kEffectiveDecimal);
  }
  {
    int uni_seed = (int)seed;
    std::default_random_engine uni_generator(uni_seed);
    std::uniform_real_distribution<double> uni_distribution(0.0, 1.0);
    for (long i = 0; i < n; ++i) rem[i] = uni_distribution(uni_generator) < ratio ? 1 : 0;
  }
}

struct Road { long dst, speed, len; };

}  // namespace

/* =============================== simulation object =============================== */
struct two_sim {
  enum Kind { PHOLD, TRAFFIC } kind = PHOLD;
  long finish_time = 0;
  /* phold */
  long num_lp = 0, num_init = 0;
  std::vector<int64_t> latency;
  std::vector<int32_t> remote;
  /* traffic */
  std::unordered_map<long, std::vector<Road> > roads;     /* state id -> out-roads */
  std::vector<long> state_ids;
  /* initial events */
  std::vector<Ev> init_events;
  /* results */
  std::vector<two_record> committed;
  uint64_t digest[4] = {0, 0, 0, 0};
  int64_t n_committed = 0;
  int64_t stats[4] = {0, 0, 0, 0};

  long initial_state(long id) const { return stateful ? (long)((uint64_t)(uint32_t)id << 32) : id; }
  bool stateful = false;   /* test model PHOLD_STATEFUL (csrc/ssb_models.cuh PholdStatefulModel); no reference app */
  std::unordered_map<long, long> final_states;

  /* event handlers: return true and fill `out` if a new event is produced; `state` may be updated */
  bool handle(const Ev& e, long& state, Ev& out) const {
    if (kind == PHOLD) {
      /* phold.hpp:256-281 */
      long ev_id = e.id;
      long src_id = e.dst;
      long send_time = e.recv;
      int num_hops = 1 + e.hops;
      int rand_table_id = (int)(ev_id + (long)num_hops) % kRandTableSize;
      uint32_t count = 0;
      if (stateful) {
        /* state = count | chain << 32; count' = count + 1; chain' = (u32) mix64(chain << 32 ^ key) */
        count = (uint32_t)((uint64_t)state & 0xFFFFFFFFull) + 1u;
        uint32_t chain = (uint32_t)((uint64_t)state >> 32);
        uint64_t key = ((uint64_t)e.recv << 32) | (uint64_t)(uint32_t)e.id;
        chain = (uint32_t)mix64(((uint64_t)chain << 32) ^ key);
        state = (long)(((uint64_t)chain << 32) | count);
        rand_table_id = (int)((ev_id + (long)num_hops + (long)(chain & 7u)) % kRandTableSize);
      }
      long receive_time = send_time + latency[rand_table_id] + kLookAhead;
      long dst_id = (remote[rand_table_id] == 1) ? latency[rand_table_id] % num_lp : e.dst;
      out = Ev();
      out.id = ev_id; out.src = src_id; out.dst = dst_id; out.recv = receive_time; out.send = send_time;
      out.hops = num_hops;
      out.rpos = stateful ? (long)count : 0;   /* p1 of the device record */
      return true;
    }
    /* traffic_sim.hpp:279-342 */
    const std::vector<long>& r = *e.route;
    long remaining = (long)r.size() - e.rpos;
    if (remaining <= 1) return false;                       /* :286-288, :295-297 */
    long next = r[e.rpos + 1];
    auto st = roads.find(e.dst);
    if (st == roads.end()) return false;                    /* LP without state: reference is UB (quirk Q13) */
    const std::vector<Road>& rd = st->second;
    size_t i = 0;
    for (; i < rd.size(); ++i) if (rd[i].dst == next) break; /* std::find, first match (:300-302) */
    if (i == rd.size()) return false;                       /* :304-306 */
    long new_arrival;
    if (rd[i].speed != 0) new_arrival = rd[i].len * 60 * 60 / rd[i].speed / 1000 + e.recv + 5; /* :314-316 */
    else new_arrival = rd[i].len * 60 * 60 + e.recv;        /* :318-319 */
    out = Ev();
    out.id = e.id; out.recv = new_arrival; out.send = e.recv; out.src = r[e.rpos]; out.dst = next;
    out.route = e.route; out.rpos = e.rpos + 1;
    return true;
  }
};

extern "C" {

void two_digest_records(const two_record* recs, int64_t n, uint64_t d[4]) {
  d[0] = d[1] = d[2] = d[3] = 0;
  for (int64_t i = 0; i < n; ++i) digest_add(d, recs[i]);
}

/* ---------------- LP-level ---------------- */
struct two_lp { LP lp; };
two_lp* two_lp_new(int64_t id) { two_lp* p = new two_lp(); p->lp.id_ = id; return p; }
void two_lp_free(two_lp* p) { delete p; }
static Ev mk(int64_t id, int64_t src, int64_t dst, int64_t recv, int64_t send, int cancel) {
  Ev e; e.id = id; e.src = src; e.dst = dst; e.recv = recv; e.send = send; e.cancel = cancel != 0; return e;
}
void two_lp_buffer(two_lp* p, int64_t id, int64_t src, int64_t dst, int64_t recv, int64_t send, int cancel) {
  p->lp.buffer(mk(id, src, dst, recv, send, cancel));
}
int64_t two_lp_flush(two_lp* p, two_record* out, int64_t cap) {
  std::vector<Ev> c;
  p->lp.flush_buf(c);
  for (int64_t i = 0; i < (int64_t)c.size() && i < cap; ++i) out[i] = to_record(c[i]);
  return (int64_t)c.size();
}
int two_lp_dequeue(two_lp* p, two_record* out) {
  Ev e;
  if (!p->lp.dequeue_event(e)) return 0;
  *out = to_record(e);
  return 1;
}
void two_lp_direct_insert(two_lp* p, int64_t id, int64_t src, int64_t dst, int64_t recv, int64_t send) {
  p->lp.direct_insert(mk(id, src, dst, recv, send, 0));
}
void two_lp_local_time(two_lp* p, int64_t out[2]) { out[0] = p->lp.local_time_.time; out[1] = p->lp.local_time_.id; }
void two_lp_set_cancel(two_lp* p, int64_t id, int64_t src, int64_t dst, int64_t recv, int64_t send) {
  p->lp.eventq_.set_cancel(mk(id, src, dst, recv, send, 0));
}
void two_lp_init_state(two_lp* p, int64_t v) { p->lp.stateq_.push_back_state(v, TS::null()); } /* logical_process.hpp:99-101 */
void two_lp_update_state(two_lp* p, int64_t v, int64_t t, int64_t i) { p->lp.stateq_.push_back_state(v, TS{t, i}); }
int64_t two_lp_get_state(two_lp* p) { return p->lp.stateq_.get_state(); }
int64_t two_lp_num_states(two_lp* p) { return (int64_t)p->lp.stateq_.state_map.size(); }
int64_t two_lp_std_out(two_lp* p, int64_t t, int64_t i, two_record* out, int64_t cap) {
  int64_t n = 0;
  p->lp.eventq_.std_out(TS{t, i}, [&](const Ev& e) { if (n < cap) out[n] = to_record(e); ++n; });
  return n;
}
void two_lp_clear(two_lp* p, int64_t t, int64_t i) {
  p->lp.eventq_.release(TS{t, i});
  p->lp.eventq_.release_cancel(TS{t, i});
  p->lp.stateq_.release(TS{t, i});
}
int64_t two_lp_size_ev(two_lp* p) { return (int64_t)p->lp.eventq_.map_.size(); }
int64_t two_lp_size_can(two_lp* p) { return (int64_t)p->lp.eventq_.can_map_.size(); }

uint64_t two_fnv_fold(const int64_t* v, int64_t n) {
  uint64_t h = 1469598103934665603ULL;
  for (int64_t i = 0; i < n; ++i) h = (h ^ (uint64_t)v[i]) * 1099511628211ULL;
  return h;
}

void two_phold_tables(int64_t seed, double lambda, double ratio, int64_t n, int64_t* lat, int32_t* rem) {
  build_phold_tables(seed, lambda, ratio, n, lat, rem);
}

/* ---------------- simulation construction ---------------- */
two_sim* two_sim_phold(int64_t num_lp, int64_t num_init, double ratio, double lambda, int64_t seed) {
  two_sim* s = new two_sim();
  s->kind = two_sim::PHOLD;
  s->finish_time = 5 * kEffectiveDecimal;                   /* phold.hpp:138 */
  s->num_lp = num_lp; s->num_init = num_init;
  s->latency.resize(kRandTableSize); s->remote.resize(kRandTableSize);
  build_phold_tables(seed, lambda, ratio, kRandTableSize, s->latency.data(), s->remote.data());
  /* phold.hpp:197-209 */
  s->init_events.reserve(num_init);
  for (long id = 0; id < num_init; ++id) {
    Ev e; e.id = id; e.src = id % num_lp; e.dst = id % num_lp;
    e.recv = s->latency[id % kRandTableSize]; e.send = e.recv; e.hops = 0;
    s->init_events.push_back(e);
  }
  for (long id = 0; id < num_lp; ++id) s->state_ids.push_back(id);   /* phold.hpp:216-225 */
  return s;
}

static std::vector<std::string> split_keep_empty(const std::string& s, char sep) {
  /* boost::split(..., is_any_of(sep)) with token_compress_off */
  std::vector<std::string> out; std::string cur;
  for (char c : s) { if (c == sep) { out.push_back(cur); cur.clear(); } else cur.push_back(c); }
  out.push_back(cur);
  return out;
}

two_sim* two_sim_traffic_files(const char* road_csv, const char* trip_csv) {
  two_sim* s = new two_sim();
  s->kind = two_sim::TRAFFIC;
  s->finish_time = 10800;                                   /* traffic_sim.hpp:240-247 */
  {
    /* traffic_reader::road_read (traffic_sim.hpp:364-408), all partitions */
    std::ifstream f(road_csv);
    if (f.fail()) { delete s; return nullptr; }
    std::string line;
    while (getline(f, line)) {
      std::vector<std::string> cp = split_keep_empty(line, ';');
      long state_id = -1;
      std::vector<Road> rds;
      for (auto& r : cp) {
        std::vector<std::string> road = split_keep_empty(r, ',');
        if (road.size() < 7) continue;
        state_id = atol(road[2].c_str());
        Road x; x.dst = atol(road[3].c_str()); x.speed = atol(road[4].c_str());
        long len = atol(road[5].c_str());
        x.len = len != 0 ? len : 1;                          /* :393-397 */
        rds.push_back(x);
      }
      s->roads[state_id] = rds;
      s->state_ids.push_back(state_id);
    }
  }
  {
    /* traffic_reader::trip_read (traffic_sim.hpp:410-451) */
    std::ifstream f(trip_csv);
    if (f.fail()) { delete s; return nullptr; }
    std::string line;
    while (getline(f, line)) {
      std::vector<std::string> v = split_keep_empty(line, ',');
      if (v.size() < 5) continue;
      Ev e; e.id = atol(v[1].c_str()); e.recv = atol(v[3].c_str()); e.send = e.recv; e.src = -1;
      auto route = std::make_shared<std::vector<long> >();
      for (size_t i = 4; i < v.size(); ++i) route->push_back(atol(v[i].c_str()));
      e.route = route; e.rpos = 0; e.dst = (*route)[0];
      s->init_events.push_back(e);
    }
  }
  return s;
}

two_sim* two_sim_traffic_arrays(int64_t n_lp, const int64_t* road_start, const int64_t* road_dst,
                                const int64_t* road_speed, const int64_t* road_len, int64_t n_trips,
                                const int64_t* trip_id, const int64_t* trip_dep, const int64_t* route_start,
                                const int64_t* route_nodes) {
  two_sim* s = new two_sim();
  s->kind = two_sim::TRAFFIC;
  s->finish_time = 10800;
  for (int64_t lp = 0; lp < n_lp; ++lp) {
    std::vector<Road> rds;
    for (int64_t j = road_start[lp]; j < road_start[lp + 1]; ++j)
      rds.push_back(Road{(long)road_dst[j], (long)road_speed[j], road_len[j] != 0 ? (long)road_len[j] : 1});
    s->roads[lp] = rds;
    s->state_ids.push_back(lp);
  }
  for (int64_t t = 0; t < n_trips; ++t) {
    Ev e; e.id = trip_id[t]; e.recv = trip_dep[t]; e.send = e.recv; e.src = -1;
    auto route = std::make_shared<std::vector<long> >(route_nodes + route_start[t], route_nodes + route_start[t + 1]);
    e.route = route; e.rpos = 0; e.dst = (*route)[0];
    s->init_events.push_back(e);
  }
  return s;
}

void two_sim_free(two_sim* s) { delete s; }
void two_sim_set_finish_time(two_sim* s, int64_t t) { s->finish_time = t; }

/* ---------------- sequential emulation of runner::loop ---------------- */
int64_t two_sim_run_timewarp(two_sim* s, int num_sched, int gsync_interval, int switch_lp_interval) {
  s->committed.clear(); s->n_committed = 0;
  s->digest[0] = s->digest[1] = s->digest[2] = s->digest[3] = 0;
  std::vector<Scheduler> scheds(num_sched);
  std::unordered_map<long, LP> lps;
  auto get_lp = [&](long id) -> LP& {
    auto it = lps.find(id);
    if (it == lps.end()) {
      LP& l = lps[id];
      l.id_ = id;
      l.sched = &scheds[id % num_sched];                     /* process_scheduler.hpp:50-53 */
      return l;
    }
    return it->second;
  };
  /* runner::init (runner.hpp:79-176): states keyed null(), events buffered at destination() */
  for (long id : s->state_ids) get_lp(id).stateq_.push_back_state(s->initial_state(id), TS::null());
  for (const Ev& e : s->init_events) get_lp(e.dst).buffer(e);

  long remote_sends = 0, gvt_rounds = 0;
  auto send_event = [&](const Ev& e) { get_lp(e.dst).buffer(e); };   /* event_communicator.hpp:90-103, one rank */

  auto run_lp_aggr = [&](LP& lp) {                                   /* runner.hpp:530-570 */
    std::vector<Ev> cancels;
    lp.flush_buf(cancels);
    for (const Ev& c : cancels) send_event(c);
    for (int i = 0; i < switch_lp_interval; ++i) {
      if (lp.local_time_ == TS::max()) break;
      Ev ev;
      if (!lp.dequeue_event(ev)) break;
      bool has_state = !lp.stateq_.state_map.empty();
      long state = has_state ? lp.stateq_.get_state() : -1;
      Ev out;
      if (!s->handle(ev, state, out)) break;                         /* :552-553 */
      lp.eventq_.set_cancel(out);                                    /* :556 */
      lp.stateq_.push_back_state(state, TS{out.send, out.id});       /* :557-558 */
      if (s->kind == two_sim::PHOLD && out.dst != ev.dst) ++remote_sends;
      send_event(out);                                               /* :566 */
    }
  };

  TS global_time = TS::zero();
  const TS finish_key{s->finish_time, 0};
  for (;;) {                                                         /* runner.hpp:350-396 */
    for (int i = 0; i < num_sched; ++i) {                            /* run_lp (runner.hpp:517-528) */
      for (int k = 0; k < gsync_interval; ++k) {
        long lp_id = scheds[i].dequeue();
        if (lp_id < 0) break;
        LP& lp = get_lp(lp_id);
        run_lp_aggr(lp);
        scheds[i].queue(lp.local_time_, lp.id_);
      }
    }
    TS local_min = TS::max();
    for (auto& sc : scheds) local_min = std::min(local_min, sc.min_locals());
    /* single process, nothing in transit: GVT = min over schedulers (global_sync.hpp:128-130) */
    if (global_time < local_min) {                                   /* gvt_communicator.hpp:41-48 */
      global_time = local_min;
      ++gvt_rounds;
      TS to = std::min(global_time, finish_key);                     /* runner.hpp:377 */
      for (auto& sc : scheds)
        for (long id : sc.active_lp_)
          get_lp(id).eventq_.std_out(to, [&](const Ev& e) {
            two_record r = to_record(e);
            s->committed.push_back(r);
            digest_add(s->digest, r);
          });
      for (auto& sc : scheds)                                        /* runner::clear (runner.hpp:572-583) */
        for (long id : sc.active_lp_) {
          LP& lp = get_lp(id);
          lp.eventq_.release(global_time);
          lp.eventq_.release_cancel(global_time);
          lp.stateq_.release(global_time);
        }
    }
    if (global_time.time >= s->finish_time) break;                   /* runner.hpp:392 */
  }
  long processed = 0, cancels = 0;
  for (auto& kv : lps) { processed += kv.second.n_events; cancels += kv.second.n_cancels; }
  s->n_committed = (int64_t)s->committed.size();
  s->stats[0] = processed; s->stats[1] = cancels; s->stats[2] = gvt_rounds; s->stats[3] = remote_sends;
  return s->n_committed;
}

/* ---------------- closure ---------------- */
int64_t two_sim_run_closure(two_sim* s, int num_threads, int keep_records) {
  s->committed.clear();
  if (num_threads < 1) num_threads = 1;
  const int64_t n = (int64_t)s->init_events.size();
  std::vector<uint64_t> dg((size_t)num_threads * 4, 0);
  std::vector<std::vector<two_record> > recs(num_threads);
  std::vector<long> remote(num_threads, 0);
  auto work = [&](int t) {
    uint64_t* d = &dg[(size_t)t * 4];
    int64_t lo = n * t / num_threads, hi = n * (t + 1) / num_threads;
    for (int64_t i = lo; i < hi; ++i) {
      Ev e = s->init_events[i];
      while (e.recv < s->finish_time) {                      /* R8: committed iff receive_time < finish_time */
        two_record r = to_record(e);
        digest_add(d, r);
        if (keep_records) recs[t].push_back(r);
        Ev out;
        long dummy = 0;
        if (!s->handle(e, dummy, out)) break;
        if (s->kind == two_sim::PHOLD && out.dst != e.dst) ++remote[t];
        e = out;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < num_threads; ++t) th.emplace_back(work, t);
  for (auto& x : th) x.join();
  s->digest[0] = s->digest[1] = s->digest[2] = s->digest[3] = 0;
  long rs = 0;
  for (int t = 0; t < num_threads; ++t) {
    s->digest[0] += dg[t * 4 + 0]; s->digest[1] += dg[t * 4 + 1];
    s->digest[2] ^= dg[t * 4 + 2]; s->digest[3] += dg[t * 4 + 3];
    rs += remote[t];
    if (keep_records) s->committed.insert(s->committed.end(), recs[t].begin(), recs[t].end());
  }
  s->n_committed = (int64_t)s->digest[0];
  s->stats[0] = s->n_committed; s->stats[1] = 0; s->stats[2] = 0; s->stats[3] = rs;
  return s->n_committed;
}

/* ---------------- plain sequential discrete-event simulation (global (receive_time,id) order) ----------------
 * The definition of correctness for any Time Warp execution: same committed records, same final states. */
int64_t two_sim_run_sequential(two_sim* s, int keep_records) {
  s->committed.clear();
  s->digest[0] = s->digest[1] = s->digest[2] = s->digest[3] = 0;
  s->final_states.clear();
  struct Item { TS key; long dst; Ev ev; };
  auto cmp = [](const Item& a, const Item& b) {
    if (b.key < a.key) return true;
    if (a.key < b.key) return false;
    return b.dst < a.dst;
  };
  std::vector<Item> heap;
  heap.reserve(s->init_events.size());
  for (const Ev& e : s->init_events) heap.push_back(Item{e.key(), e.dst, e});
  std::make_heap(heap.begin(), heap.end(), cmp);
  for (long id : s->state_ids) s->final_states[id] = s->initial_state(id);
  long processed = 0;
  while (!heap.empty()) {
    std::pop_heap(heap.begin(), heap.end(), cmp);
    Item it = heap.back();
    heap.pop_back();
    if (it.ev.recv >= s->finish_time) continue;
    two_record r = to_record(it.ev);
    digest_add(s->digest, r);
    if (keep_records) s->committed.push_back(r);
    ++processed;
    long& st = s->final_states[it.ev.dst];
    long st_new = st;
    Ev out;
    if (!s->handle(it.ev, st_new, out)) continue;   /* no event: state update dropped (runner.hpp:552-558) */
    st = st_new;
    heap.push_back(Item{out.key(), out.dst, out});
    std::push_heap(heap.begin(), heap.end(), cmp);
  }
  s->n_committed = (int64_t)s->digest[0];
  s->stats[0] = processed; s->stats[1] = 0; s->stats[2] = 0; s->stats[3] = 0;
  return s->n_committed;
}

void two_sim_set_stateful(two_sim* s, int on) { s->stateful = on != 0; }

/* final LP states of the last run_sequential, as the device's two state words {lo, hi} */
int64_t two_sim_final_states(two_sim* s, int64_t* ids, uint32_t* words, int64_t cap) {
  int64_t n = 0;
  std::vector<long> keys;
  for (auto& kv : s->final_states) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  for (long id : keys) {
    if (n >= cap) break;
    uint64_t v = (uint64_t)s->final_states[id];
    ids[n] = id; words[2 * n] = (uint32_t)v; words[2 * n + 1] = (uint32_t)(v >> 32);
    ++n;
  }
  return n;
}

int64_t two_sim_num_committed(two_sim* s) { return s->n_committed; }
int64_t two_sim_committed(two_sim* s, two_record* out, int64_t cap) {
  int64_t n = std::min<int64_t>(cap, (int64_t)s->committed.size());
  std::copy(s->committed.begin(), s->committed.begin() + n, out);
  return n;
}
void two_sim_digest(two_sim* s, uint64_t d[4]) { for (int i = 0; i < 4; ++i) d[i] = s->digest[i]; }
void two_sim_stats(two_sim* s, int64_t st[4]) { for (int i = 0; i < 4; ++i) st[i] = s->stats[i]; }

}  /* extern "C" */
