// scalesim::runner<App> — same entry points as the reference's driver (include/scalesim/simulation/runner.hpp:28-60:
// init_main(&argc, &argv), run()), but the Time Warp core behind it is the device engine of libscalesim_b200.so:
//
//   run():  App::init()                                      (runner.hpp:82)
//           App::init_partition_index(world)  -> ssb_set_partition            (runner.hpp:98)
//           App::init_states_in_this_rank     -> ssb_load_states              (runner.hpp:108-118)
//           App::init_events + shuffle by partition[destination] -> ssb_load_events   (runner.hpp:131-147)
//           ssb_run_stream_packed             = runner::loop (runner.hpp:350-396): the window loop runs as one CUDA
//                                               graph while the committed records stream out through pinned staging
//                                               buffers and are printed as "RE,..." lines (queue.hpp:203-211,
//                                               sim_obj.hpp:66-77) — the reference prints while it simulates, too
//           --diff_init: ssb_run_stream_history — the same with every event's sent-log entry; a writer thread appends
//                                               the chunks to the history store (logical_process.hpp:187-203)
//           report on stderr                                                    (runner.hpp:482-507)
//
// The application's host event_handler is not on the hot path: its device twin runs in the kernels. It is still
// used here as the checker of that twin: with SCALESIM_B200_VERIFY=<k> the chains of the first k initial events are
// also advanced on the host with App::event_handler and must be among the committed records.
//
// One process drives one GPU (one partition). Several partitions: start `world` processes with
// SCALESIM_B200_RANK / SCALESIM_B200_WORLD set and SCALESIM_B200_COMM_FILE pointing to a file they can all reach
// (rank 0 writes the NCCL unique id there); CUDA device = SCALESIM_B200_DEVICE or the rank.
// Tunables: SCALESIM_B200_BUCKET (ticks), SCALESIM_B200_WINDOW (buckets), SCALESIM_B200_MAX_EVENTS,
// SCALESIM_B200_MAX_BATCH, SCALESIM_B200_QUIET=1 (digest instead of result lines), SCALESIM_B200_P2P=0 (exchange by
// ncclSend/ncclRecv instead of peer-memory writes).
// Input at scale: SCALESIM_B200_SCENARIO_OUT=<file> (one partition) writes what the application's readers / generators
// produced — partition table, packed states, model data, initial events as device records — as a scenario image
// (ssb_scenario_save); SCALESIM_B200_SCENARIO=<file> starts from such an image instead (ssb_scenario_open +
// ssb_scenario_load: one mapping, no parse, no per-event object; any rank / world), e.g. one made from CSV files by
// ssb_traffic_compile. The application's init() still runs; its init_partition_index / init_states_in_this_rank /
// init_events do not.
#ifndef SCALESIM_B200_SIMULATION_RUNNER_HPP_
#define SCALESIM_B200_SIMULATION_RUNNER_HPP_

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <filesystem>
#include <fstream>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>
#include <boost/make_shared.hpp>
#include <boost/optional.hpp>
#include <boost/shared_ptr.hpp>
#include <gflags/gflags.h>
#include <glog/logging.h>
#include "scalesim/simulation/device_model.hpp"
#include "scalesim/util.hpp"
#include "scalesim_b200.h"

namespace scalesim {

template <class App>
class runner {
 public:
  runner() : engine_(NULL), rank_(0), world_(1) {}
  virtual ~runner() { if (engine_) ssb_destroy(engine_); }
  void init_main(int* argc, char** argv[]);
  void run();

 private:
  runner(const runner&);
  void operator=(const runner&);
  typedef device_model<App> model;
  App app_;
  ssb_engine* engine_;
  int rank_, world_;

  static long env_long(const char* name, long dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atol(v) : dflt;
  }
  void fail(const std::string& what, int rc) {
    LOG(ERROR) << what << " failed (status " << rc << "): " << ssb_last_error(engine_);
    std::exit(1);
  }
  void comm_init();
  std::vector<std::string> comm_files_;      // rendezvous files this rank wrote (removed at the end of run())
  // exact-differential history store of this rank: <store_dir>/<sim_id>/history<rank>.ssb (same file format as
  // scalesim_b200/store.py): the committed events with their sent-log entries, the LPs that had an initial state and
  // the model data the events refer to (TrafficSim: the route pool)
  struct history {
    std::vector<ssb_raw_record> rec;
    std::vector<ssb_sent_record> sent;
    std::vector<int64_t> state_lps, pool;
  };
  // Appends (records, sent-log entries) chunks to the history file from a background thread while the simulation
  // runs: the reference keeps every put in one in-memory WriteBatch until finish() (leveldb_store.hpp:132-141, quirk
  // Q15); here host memory is bounded by the queue depth. File = "SSBHIST3", n_state, n_pool, state LPs, model data
  // pool, then chunks {n, rec[n], sent[n]} closed by n = 0 and the total (same format as scalesim_b200/store.py).
  class history_writer {
   public:
    history_writer() : f_(NULL), total_(0), bytes_(0), stop_(false), failed_(false) {}
    bool open(const std::string& path, const std::vector<int64_t>& state_lps, const std::vector<int64_t>& pool) {
      path_ = path;
      f_ = std::fopen((path + ".tmp").c_str(), "wb");
      if (!f_) return false;
      const int64_t hdr[2] = {(int64_t)state_lps.size(), (int64_t)pool.size()};
      put("SSBHIST3", 8); put(hdr, sizeof(hdr));
      put(state_lps.data(), state_lps.size() * 8); put(pool.data(), pool.size() * 8);
      thr_ = std::thread(&history_writer::loop, this);
      return !failed_;
    }
    // called from the drain sink: the staging buffer is reused at once, so the chunk is copied into the queue
    void push(const ssb_raw_record* rec, const ssb_sent_record* sent, int64_t n) {
      chunk c;
      c.rec.assign(rec, rec + n); c.sent.assign(sent, sent + n);
      std::unique_lock<std::mutex> lk(m_);
      cv_.wait(lk, [&] { return q_.size() < 4; });      // back-pressure: at most 4 chunks in flight
      q_.push_back(std::move(c));
      cv_.notify_all();
    }
    bool close() {
      { std::unique_lock<std::mutex> lk(m_); stop_ = true; cv_.notify_all(); }
      if (thr_.joinable()) thr_.join();
      if (!f_) return false;
      const int64_t tail[2] = {0, total_};
      put(tail, sizeof(tail));
      const bool ok = std::fclose(f_) == 0 && !failed_;
      f_ = NULL;
      return ok && std::rename((path_ + ".tmp").c_str(), path_.c_str()) == 0;
    }
    int64_t records() const { return total_; }
    int64_t bytes() const { return bytes_; }
   private:
    struct chunk { std::vector<ssb_raw_record> rec; std::vector<ssb_sent_record> sent; };
    void put(const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f_) != n) failed_ = true; bytes_ += (int64_t)n; }
    void loop() {
      for (;;) {
        chunk c;
        {
          std::unique_lock<std::mutex> lk(m_);
          cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
          if (q_.empty()) return;
          c = std::move(q_.front());
          q_.pop_front();
          cv_.notify_all();
        }
        const int64_t n = (int64_t)c.rec.size();
        put(&n, 8); put(c.rec.data(), c.rec.size() * sizeof(ssb_raw_record)); put(c.sent.data(), c.sent.size() * sizeof(ssb_sent_record));
        total_ += n;
      }
    }
    std::FILE* f_;
    std::string path_;
    std::thread thr_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<chunk> q_;
    int64_t total_, bytes_;
    bool stop_, failed_;
  };
  // "RE,id,src,send,dst,recv[,START|,END]\n" (sim_obj.hpp:66-77) for a chunk of records, formatted by a few threads
  // into per-thread buffers and written in order
  struct line_printer {
    static char* num(char* p, uint32_t v) {
      char t[10]; int n = 0;
      do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
      while (n) *p++ = t[--n];
      return p;
    }
    static char* line(char* p, uint32_t id, uint32_t src, uint32_t send, uint32_t dst, uint32_t recv, uint32_t tag) {
      *p++ = 'R'; *p++ = 'E'; *p++ = ',';
      p = num(p, id); *p++ = ',';
      if (src == 0xffffffffu) { *p++ = '-'; *p++ = '1'; } else p = num(p, src);
      *p++ = ','; p = num(p, send); *p++ = ','; p = num(p, dst); *p++ = ','; p = num(p, recv);
      if (tag == 1) { std::memcpy(p, ",START", 6); p += 6; } else if (tag == 2) { std::memcpy(p, ",END", 4); p += 4; }
      *p++ = '\n';
      return p;
    }
    std::vector<std::vector<char> > bufs;
    template <class Get> void print(int64_t n, Get get) {
      const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(8, std::min<int64_t>(n / 65536 + 1, std::thread::hardware_concurrency())));
      bufs.resize((size_t)nt);
      std::vector<size_t> used((size_t)nt, 0);
      auto work = [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        std::vector<char>& buf = bufs[(size_t)t];
        buf.resize((size_t)(b - a) * 64 + 64);
        char* p = buf.data();
        for (int64_t i = a; i < b; ++i) p = get(i, p);
        used[(size_t)t] = (size_t)(p - buf.data());
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
      work(0);
      for (size_t t = 0; t < th.size(); ++t) th[t].join();
      for (int t = 0; t < nt; ++t) std::fwrite(bufs[(size_t)t].data(), 1, used[(size_t)t], stdout);
    }
  };
  std::string history_path() const {
    return (FLAGS_store_dir.empty() ? std::string(".") : FLAGS_store_dir) + "/" + FLAGS_sim_id + "/history" + std::to_string(rank_) + ".ssb";
  }
  void read_history(history& h);
  void write_image(const char* path, const std::vector<long>& part, const std::vector<int64_t>& state_lps,
                   const std::vector<uint32_t>& words, const std::vector<ssb_event>& pod, typename model::context& cx, long n_lp);
  void verify_twin(const ev_vec<App>& initial, const std::vector<ssb_record>& committed, long k);
};

template <class App>
void runner<App>::init_main(int* argc, char** argv[]) {
  gflags::SetUsageMessage(
      "\n=====How to run simulation code====\n"
      "    $ {Binary} {Application Args}\n"
      "Example of TrafficSim:\n"
      "    $ ./apps/TrafficSim {partition file} {road file} {trip file}\n"
      "Example of PHOLD:\n"
      "    $ ./apps/PHOLD {# of LPs} {# of initial messages} {remote ratio} {lambda} {# of what-ifs} {cut interval}\n"
      "Several GPUs: one process per GPU with SCALESIM_B200_RANK / SCALESIM_B200_WORLD / SCALESIM_B200_COMM_FILE.\n"
      "===================================");
  gflags::ParseCommandLineFlags(argc, argv, true);
  FLAGS_logtostderr = 1;
  FLAGS_minloglevel = 0;
  google::InitGoogleLogging((*argv)[0]);
  google::InstallFailureSignalHandler();
  rank_ = (int)env_long("SCALESIM_B200_RANK", 0);
  world_ = (int)env_long("SCALESIM_B200_WORLD", 1);
}

template <class App>
void runner<App>::comm_init() {
  if (world_ <= 1) return;
  const char* path = std::getenv("SCALESIM_B200_COMM_FILE");
  if (!path) { LOG(ERROR) << "SCALESIM_B200_COMM_FILE must be set when SCALESIM_B200_WORLD > 1"; std::exit(1); }
  const bool p2p = env_long("SCALESIM_B200_P2P", 1) != 0;
  // Rendezvous through files next to COMM_FILE. Every file starts with the id of THIS launch (SCALESIM_B200_RUN_ID,
  // default: the parent process id, which the ranks of one launch share), so that a file left behind by an earlier
  // run with the same COMM_FILE is never taken for a peer's: a reader polls until the id matches. Each rank removes
  // its own file when it is done.
  const uint64_t run_id = (uint64_t)env_long("SCALESIM_B200_RUN_ID", (long)getppid());
  auto publish = [&](const std::string& f, const char* data, size_t n) {
    std::string tmp = f + ".tmp" + std::to_string(rank_);
    { std::ofstream o(tmp.c_str(), std::ios::binary); o.write((const char*)&run_id, 8); o.write(data, (std::streamsize)n); }
    std::rename(tmp.c_str(), f.c_str());
    comm_files_.push_back(f);
  };
  auto fetch = [&](const std::string& f, char* data, size_t n) {
    for (int tries = 0;; ++tries) {
      std::ifstream in(f.c_str(), std::ios::binary);
      uint64_t id = 0;
      if (in && in.read((char*)&id, 8) && id == run_id && in.read(data, (std::streamsize)n) && (size_t)in.gcount() == n) return;
      if (tries > 600) { LOG(ERROR) << "timed out waiting for " << f << " (run id " << run_id << ")"; std::exit(1); }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
  };
  int rc = 0;
  if (!p2p) {      // events travel by ncclSend/ncclRecv, GVT by ncclAllReduce: the ranks join one NCCL communicator
    char id[128];
    if (rank_ == 0) {
      if ((rc = ssb_comm_unique_id(id))) fail("ssb_comm_unique_id", rc);
      publish(path, id, 128);
    } else {
      fetch(path, id, 128);
    }
    if ((rc = ssb_comm_init(engine_, id))) fail("ssb_comm_init", rc);
    return;
  }
  // peer-memory exchange (default; events AND the GVT reduction cross through the mailboxes, no NCCL): swap the CUDA
  // IPC handles of the engines' mailboxes
  char mine[64];
  if ((rc = ssb_comm_export(engine_, mine))) fail("ssb_comm_export", rc);
  publish(std::string(path) + ".h" + std::to_string(rank_), mine, 64);
  std::vector<char> all((size_t)world_ * 64);
  for (int r = 0; r < world_; ++r) fetch(std::string(path) + ".h" + std::to_string(r), &all[(size_t)r * 64], 64);
  if ((rc = ssb_comm_import(engine_, all.data(), world_))) fail("ssb_comm_import", rc);
}

template <class App>
void runner<App>::read_history(history& h) {
  const std::string p = history_path();
  std::ifstream f(p.c_str(), std::ios::binary);
  char magic[8];
  int64_t hdr[2];
  if (!f || !f.read(magic, 8) || std::memcmp(magic, "SSBHIST3", 8) != 0 || !f.read((char*)hdr, sizeof(hdr))) {
    LOG(ERROR) << "no history store at " << p << " (run with --diff_init first)";
    std::exit(1);
  }
  std::error_code ec;
  const int64_t fsize = (int64_t)std::filesystem::file_size(p, ec);
  auto bad = [&](const char* why) { LOG(ERROR) << p << ": " << why; std::exit(1); };
  if (hdr[0] < 0 || hdr[1] < 0 || (hdr[0] + hdr[1]) * 8 > fsize) bad("header counts do not fit the file");
  h.state_lps.resize((size_t)hdr[0]); h.pool.resize((size_t)hdr[1]);
  f.read((char*)h.state_lps.data(), (std::streamsize)(h.state_lps.size() * 8));
  f.read((char*)h.pool.data(), (std::streamsize)(h.pool.size() * 8));
  for (;;) {
    int64_t n = -1;
    if (!f.read((char*)&n, 8)) bad("truncated (no closing chunk)");
    if (n == 0) break;
    if (n < 0 || n * 48 > fsize) bad("bad chunk size");
    const size_t at = h.rec.size();
    h.rec.resize(at + (size_t)n); h.sent.resize(at + (size_t)n);
    f.read((char*)&h.rec[at], (std::streamsize)((size_t)n * sizeof(ssb_raw_record)));
    f.read((char*)&h.sent[at], (std::streamsize)((size_t)n * sizeof(ssb_sent_record)));
    if (!f) bad("truncated chunk");
  }
  int64_t total = -1;
  if (!f.read((char*)&total, 8) || total != (int64_t)h.rec.size()) bad("record count does not match the chunks");
}

// SCALESIM_B200_SCENARIO_OUT: what the application's readers / generators produced — partition table, packed states, model
// data, initial events as device records — kept as a scenario image (ssb_scenario_save) for later runs
template <class App>
void runner<App>::write_image(const char* path, const std::vector<long>& part, const std::vector<int64_t>& state_lps,
                              const std::vector<uint32_t>& words, const std::vector<ssb_event>& pod, typename model::context& cx,
                              long n_lp) {
  if (world_ != 1) { LOG(ERROR) << "SCALESIM_B200_SCENARIO_OUT needs a single partition (the image holds every partition's share)"; std::exit(1); }
  std::vector<ssb_raw_record> raw(pod.size());
  for (size_t i = 0; i < pod.size(); ++i) {
    const ssb_event& e = pod[i];
    if (e.recv_time < 0 || e.recv_time >= (1ll << 31) || e.send_time < 0 || e.send_time >= (1ll << 31) || e.id < 0 || e.id >= 0xffffffffll ||
        e.dst < 0 || e.dst >= 0xffffffffll || e.src < -1 || e.src >= 0xffffffffll || e.p0 < 0 || e.p0 > 0xffffffffll || e.p1 < 0 || e.p1 > 0xffffffffll) {
      LOG(ERROR) << "event " << e.id << " does not fit the device record (see the limits in scalesim_b200.h)"; std::exit(1);
    }
    ssb_raw_record& r = raw[i];
    r.key = ((uint64_t)e.recv_time << 32) | (uint64_t)e.id;
    r.dst = (uint32_t)e.dst; r.src = e.src < 0 ? 0xffffffffu : (uint32_t)e.src; r.send_time = (uint32_t)e.send_time;
    r.p0 = (uint32_t)e.p0; r.p1 = (uint32_t)e.p1; r.tag = (uint32_t)(e.flags & 1);
  }
  std::vector<int64_t> part_raw(part.begin(), part.end());
  ssb_scenario out;
  std::memset(&out, 0, sizeof(out));
  out.model = model::model_id(); out.state_words = model::state_words(); out.n_lp = n_lp; out.finish_time = App::finish_time();
  out.n_part = (int64_t)part_raw.size(); out.lp_to_part = part_raw.data();
  out.n_states = (int64_t)state_lps.size(); out.state_lps = state_lps.data(); out.states = words.data();
  out.n_events = (int64_t)raw.size(); out.events = raw.data();
  model::image_slots(cx, out);
  int rc = ssb_scenario_save(path, &out);
  if (rc) { LOG(ERROR) << "ssb_scenario_save failed (status " << rc << "): " << ssb_last_error(NULL); std::exit(1); }
  LOG(INFO) << "Scenario image written: " << path;
}

template <class App>
void runner<App>::run() {
  if (FLAGS_diff_init && FLAGS_diff_repeat) { LOG(ERROR) << "--diff_init and --diff_repeat exclude each other"; std::exit(2); }
  typedef std::chrono::steady_clock clk;
  clk::time_point t0 = clk::now();
  app_.init();
  LOG_IF(INFO, rank_ == 0) << "Initiate application";

  // a scenario image replaces the application's readers (not for --diff_repeat: that run starts from the history store)
  const char* image_path = std::getenv("SCALESIM_B200_SCENARIO");
  const bool from_image = image_path && *image_path && !FLAGS_diff_repeat;
  ssb_scenario img;
  std::memset(&img, 0, sizeof(img));
  if (from_image) {
    int irc = ssb_scenario_open(image_path, &img);
    if (irc) { LOG(ERROR) << "ssb_scenario_open failed (status " << irc << "): " << ssb_last_error(NULL); std::exit(1); }
    if (img.model != model::model_id()) { LOG(ERROR) << image_path << " is a scenario of another application (model " << img.model << ")"; std::exit(1); }
    if (img.finish_time != App::finish_time())
      LOG(WARNING) << image_path << ": finish time " << img.finish_time << " differs from the application's " << App::finish_time() << "; the application's is used";
    LOG_IF(INFO, rank_ == 0) << "Scenario image " << image_path << ": " << img.n_lp << " LPs, " << img.n_states << " states, " << img.n_events << " events";
  }
  auto image_owner = [&](long lp) -> int {
    return img.n_part ? (lp < img.n_part ? (int)(img.lp_to_part[lp] % world_) : 0) : (int)(lp % world_);
  };

  std::pair<parti_ptr, parti_indx_ptr> parti;
  if (!from_image) parti = app_.init_partition_index(world_);
  static const std::vector<long> no_part;
  const std::vector<long>& part = from_image ? no_part : *parti.first;
  const long n_lp = from_image ? (long)img.n_lp : model::n_lp(parti.first);
  typename model::context cx;

  // states of this rank
  st_vec<App> states;
  if (!from_image) app_.init_states_in_this_rank(states, rank_, world_, parti.first);
  // events: every rank reads the whole scenario and keeps what it owns (the reference shuffles instead, runner.hpp:131-147)
  ev_vec<App> mine;
  history hist;
  std::vector<ssb_event> deletes;      // anti-messages of the DE / RE what-if queries
  struct state_change { long lp, time; std::vector<uint32_t> words; };
  std::vector<state_change> state_changes;      // SC what-if queries
  if (from_image) {
    // nothing to read: the image holds the events as device records
  } else if (!FLAGS_diff_repeat) {
    // ONE pass over the scenario: the application shards its input by "index % rank_size == rank" and "any way of
    // reading is fine, the events are shuffled to their LPs afterwards" (phold.hpp:190-209, traffic_sim.hpp:410-451),
    // so asking for shard 0 of 1 yields every event; this rank keeps the ones whose LP it owns (the reference shuffles
    // instead, runner.hpp:131-147). No rank parses the input `world` times.
    ev_vec<App> parsed;
    app_.init_events(parsed, 0, 1);
    for (typename ev_vec<App>::iterator it = parsed.begin(); it != parsed.end(); ++it) {
      long d = (*it)->destination();
      if (d >= 0 && d < n_lp && part[d] % world_ == rank_) mine.push_back(*it);
    }
  } else {
    // --diff_repeat (runner::init_repeat, runner.hpp:178-348): no events are read from the scenario. The recorded
    // history of this rank comes back from the store, the what-if queries (App::init_what_if) are injected on top.
    read_history(hist);
    model::restore_model_data(cx, hist.pool);
    {
      std::vector<boost::shared_ptr<const what_if<App> > > wi;
      std::vector<std::pair<uint64_t, uint32_t> > by_key;      // (key, index) of the recorded events, sorted; built on demand
      app_.init_what_if(wi, 0, 1);      // every query in one pass (same sharding contract as init_events)
      for (size_t i = 0; i < wi.size(); ++i) {
        long d = wi[i]->lp_id_;
        if (!(d >= 0 && d < n_lp && part[d] % world_ == rank_)) continue;
        if (wi[i]->update_state_.id() != -1) {
          // state change (runner.hpp:217-245): the LP gets this state at (time, 0); its recorded events and sent
          // events from there on are re-armed -> ssb_set_state_from after the history is loaded
          state_change sc;
          sc.lp = d; sc.time = wi[i]->time_;
          sc.words.resize((size_t)model::state_words());
          model::pack_state(wi[i]->update_state_, sc.words.data(), cx);
          state_changes.push_back(sc);
        }
        if (wi[i]->delete_event_id_ != -1) {
          // delete event (runner.hpp:246-276): the recorded event loses its sent-log entry — eventq::delete_can: its
          // output is NOT cancelled — and an anti-message goes into the LP's inbox (eventq::delete_ev,
          // queue.hpp:227-241), which erases the event and rolls the LP back to its key
          const uint64_t key = ((uint64_t)wi[i]->time_ << 32) | (uint64_t)(uint32_t)wi[i]->delete_event_id_;
          if (by_key.empty() && !hist.rec.empty()) {      // first delete query: index the history once (key -> record)
            by_key.resize(hist.rec.size());
            for (size_t k = 0; k < by_key.size(); ++k) by_key[k] = std::make_pair(hist.rec[k].key, (uint32_t)k);
            std::sort(by_key.begin(), by_key.end());
          }
          for (std::vector<std::pair<uint64_t, uint32_t> >::const_iterator it =
                   std::lower_bound(by_key.begin(), by_key.end(), std::make_pair(key, (uint32_t)0));
               it != by_key.end() && it->first == key; ++it)
            if ((long)hist.rec[it->second].dst == d) hist.sent[it->second].dst = 0xffffffffu;
          ssb_event a;
          std::memset(&a, 0, sizeof(a));
          a.id = wi[i]->delete_event_id_; a.src = -1; a.dst = d; a.recv_time = a.send_time = wi[i]->time_; a.flags = 1;
          deletes.push_back(a);
        }
        if (wi[i]->add_event_.id() != -1) mine.push_back(boost::make_shared<event<App> >(wi[i]->add_event_));
      }
    }
  }
  std::vector<ssb_event> pod(mine.size());
  for (size_t i = 0; i < mine.size(); ++i) model::pack_event(*mine[i], pod[i], cx);
  pod.insert(pod.end(), deletes.begin(), deletes.end());
  // from an image: the events of this rank's LPs are counted (and the committed log sized) in one pass over the mapping
  size_t n_image_events = 0, n_image_states = 0;
  long image_committed_bound = 0;
  if (from_image) {
    for (int64_t i = 0; i < img.n_events; ++i) {
      if (world_ > 1 && image_owner((long)img.events[i].dst) != rank_) continue;
      ++n_image_events;
      image_committed_bound += model::committed_bound_one(img.events[i], App::finish_time());
    }
    for (int64_t i = 0; i < img.n_states; ++i)
      if (world_ == 1 || image_owner((long)img.state_lps[i]) == rank_) { ++n_image_states; if (FLAGS_diff_init) hist.state_lps.push_back(img.state_lps[i]); }
  }
  const size_t n_resident = pod.size() + hist.rec.size() + n_image_events;

  ssb_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = SSB_ABI_VERSION;
  cfg.model = model::model_id();
  cfg.device = (int)env_long("SCALESIM_B200_DEVICE", rank_);
  cfg.rank = rank_; cfg.world = world_;
  cfg.n_lp = n_lp;
  cfg.finish_time = App::finish_time();
  cfg.bucket_ticks = env_long("SCALESIM_B200_BUCKET", model::default_bucket_ticks());
  cfg.window_buckets = (int)env_long("SCALESIM_B200_WINDOW", 1);
  // a repeat run holds the recorded copy AND the re-sent copy of what it re-executes until the bucket is fossil-collected
  cfg.max_events = env_long("SCALESIM_B200_MAX_EVENTS", (long)(n_resident * (FLAGS_diff_repeat ? 3.0 : 1.6)) + (1 << 18));
  cfg.max_batch = env_long("SCALESIM_B200_MAX_BATCH", (long)n_resident / 2 + (1 << 16));
  const bool quiet = env_long("SCALESIM_B200_QUIET", 0) != 0 && !FLAGS_diff_init;
  // the device log holds the whole run (it streams out while the loop runs, but the device is never throttled by the
  // host): sized from the workload; an overflow is a loud SSB_ERR_CAPACITY, never a silently shorter output
  cfg.max_committed = quiet ? 0 : env_long("SCALESIM_B200_MAX_COMMITTED",
                                           (long)hist.rec.size() + model::committed_bound(pod, App::finish_time()) + image_committed_bound + (1 << 16));
  if (FLAGS_diff_init) cfg.flags |= SSB_FLAG_HISTORY;
  cfg.max_snapshots = 1 << 20;
  cfg.rounds_per_sync = 4;
  int rc = ssb_create(&cfg, &engine_);
  if (rc) { LOG(ERROR) << "ssb_create failed (status " << rc << "): " << ssb_last_error(NULL); std::exit(1); }
  comm_init();
  const char* image_out = std::getenv("SCALESIM_B200_SCENARIO_OUT");
  if (from_image) {
    // partition, model data, this rank's states and events straight from the mapping
    if ((rc = ssb_scenario_load(engine_, &img))) fail("ssb_scenario_load", rc);
    if (FLAGS_diff_init) hist.pool = model::model_data_blob_of(img);
    ssb_scenario_close(&img);
  } else {
    std::vector<int64_t> part64(part.begin(), part.end());
    for (size_t i = 0; i < part64.size(); ++i) part64[i] %= world_;
    if ((rc = ssb_set_partition(engine_, part64.data(), (int64_t)part64.size()))) fail("ssb_set_partition", rc);
    const int sw = model::state_words();
    std::vector<int64_t> ids(states.size());
    std::vector<uint32_t> words(states.size() * (size_t)sw);
    for (size_t i = 0; i < states.size(); ++i) {
      ids[i] = states[i]->id();
      model::pack_state(*states[i], &words[i * (size_t)sw], cx);      // may add to the model data (TrafficSim road table)
    }
    if ((rc = model::upload_model_data(engine_, cx))) fail("ssb_set_model_data", rc);
    if ((rc = ssb_load_states(engine_, ids.data(), words.data(), (int64_t)ids.size()))) fail("ssb_load_states", rc);
    if (image_out && *image_out && !FLAGS_diff_repeat) write_image(image_out, part, ids, words, pod, cx, n_lp);
  }
  if (FLAGS_diff_repeat &&
      (rc = ssb_load_history(engine_, hist.rec.data(), hist.sent.data(), (int64_t)hist.rec.size()))) fail("ssb_load_history", rc);
  for (size_t i = 0; i < state_changes.size(); ++i)
    if ((rc = ssb_set_state_from(engine_, state_changes[i].lp, state_changes[i].time, state_changes[i].words.data()))) fail("ssb_set_state_from", rc);
  if ((rc = ssb_load_events(engine_, pod.data(), (int64_t)pod.size()))) fail("ssb_load_events", rc);
  double init_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  LOG_IF(INFO, rank_ == 0)
      << "\n====================================================================\n"
      << " Number of partitions: " << world_ << " (one B200 each)\n"
      << "     Total events num (this partition): " << pod.size() + n_image_events << "\n"
      << "     Total states num (this partition): " << states.size() + n_image_states << "\n"
      << " Total init time: " << (long)init_ms << " ms\n"
      << "====================================================================\n";

  LOG_IF(INFO, rank_ == 0) << "Start Simulation";
  std::vector<ssb_record> kept;
  const long verify_k = env_long("SCALESIM_B200_VERIFY", 0);
  if (verify_k > 0 && from_image) LOG(INFO) << "SCALESIM_B200_VERIFY walks the application's own initial events; skipped for a scenario image";
  const bool keep = verify_k > 0 && !quiet && !FLAGS_diff_init && !FLAGS_diff_repeat && !from_image;
  clk::time_point t1 = clk::now();
  ssb_stats st;
  std::memset(&st, 0, sizeof(st));
  std::setvbuf(stdout, NULL, _IOFBF, 1 << 22);
  line_printer lp;
  history_writer hw;
  double print_ms = 0.0;
  struct sink_ctx { runner* self; line_printer* lp; history_writer* hw; std::vector<ssb_record>* kept; bool keep; double* print_ms; };
  sink_ctx sc = {this, &lp, &hw, &kept, keep, &print_ms};
  if (FLAGS_diff_init) {
    // --diff_init: the committed events leave the device together with their sent-log entries while the loop runs;
    // the lines are printed from the same records, the store file grows on a writer thread
    for (size_t i = 0; i < states.size(); ++i) hist.state_lps.push_back(states[i]->id());
    if (!from_image) hist.pool = model::model_data_blob(cx);
    const std::string p = history_path(), dir = p.substr(0, p.find_last_of('/'));
    std::error_code ec;
    std::filesystem::create_directories(dir, ec);
    if (ec || !hw.open(p, hist.state_lps, hist.pool)) { LOG(ERROR) << "cannot write " << p; std::exit(1); }
    struct fn {
      static int sink(void* u, const ssb_raw_record* rec, const ssb_sent_record* sent, int64_t n) {
        sink_ctx* c = (sink_ctx*)u;
        clk::time_point a = clk::now();
        c->lp->print(n, [&](int64_t i, char* p) {
          const ssb_raw_record& r = rec[i];
          return line_printer::line(p, (uint32_t)(r.key & 0xffffffffu), r.src, r.send_time, r.dst, (uint32_t)(r.key >> 32), r.tag);
        });
        c->hw->push(rec, sent, n);
        *c->print_ms += std::chrono::duration<double, std::milli>(clk::now() - a).count();
        return 0;
      }
    };
    if ((rc = ssb_run_stream_history(engine_, &fn::sink, &sc, &st))) fail("ssb_run_stream_history", rc);
  } else if (!quiet) {
    struct fn {
      static int sink(void* u, const ssb_packed_record* rec, int64_t n) {
        sink_ctx* c = (sink_ctx*)u;
        clk::time_point a = clk::now();
        c->lp->print(n, [&](int64_t i, char* p) {
          const ssb_packed_record& r = rec[i];
          return line_printer::line(p, r.id, r.src, r.send_time, r.dst, r.recv_time, r.tag);
        });
        if (c->keep) {
          for (int64_t i = 0; i < n; ++i) {
            const ssb_packed_record& r = rec[i];
            ssb_record w = {(int64_t)r.id, r.src == 0xffffffffu ? -1 : (int64_t)r.src, (int64_t)r.send_time, (int64_t)r.dst,
                            (int64_t)r.recv_time, (int64_t)r.tag};
            c->kept->push_back(w);
          }
        }
        *c->print_ms += std::chrono::duration<double, std::milli>(clk::now() - a).count();
        return 0;
      }
    };
    if ((rc = ssb_run_stream_packed(engine_, &fn::sink, &sc, &st))) fail("ssb_run_stream_packed", rc);
  } else {
    if ((rc = ssb_run(engine_, &st))) fail("ssb_run", rc);
  }
  std::fflush(stdout);
  double sim_ms = std::chrono::duration<double, std::milli>(clk::now() - t1).count();
  uint64_t dg[4];
  ssb_committed_digest(engine_, dg);
  if (keep) verify_twin(mine, kept, verify_k);
  double store_ms = 0.0;
  if (FLAGS_diff_init) {
    clk::time_point a = clk::now();
    if (!hw.close()) { LOG(ERROR) << "cannot write " << history_path(); std::exit(1); }
    store_ms = std::chrono::duration<double, std::milli>(clk::now() - a).count();
    LOG(INFO) << " Stored objects size : " << hw.bytes() << " byte (" << hw.records() << " events with their sent-log entries, "
              << hist.state_lps.size() << " initial states) -> " << history_path();
  }

  LOG(INFO) << "Finish Simulation";
  LOG(INFO)
      << "\n====================================================================\n"
      << " Partition " << rank_ << " of " << world_ << "\n"
      << " Simulation elapsed time: " << (long)sim_ms << " ms (window loop + streaming out " << (quiet ? "nothing" : "and printing the result lines") << ")\n"
      << "   Window loop on the device        : " << st.loop_ms << " ms\n"
      << "   Init (parse, pack, load)         : " << (long)init_ms << " ms\n"
      << "   OutPutResult (format + print)    : " << (long)print_ms << " ms (inside the elapsed time, overlapped with the loop)\n"
      << "   DBWrite (flush the history store): " << (long)store_ms << " ms\n"
      << "\n"
      << " # of global synchronizations   : " << st.rounds << "\n"
      << " # of processing events         : " << st.processed << "\n"
      << " # of cancels (= # of rollback) : " << st.antis_sent << "\n"
      << " # of output events             : " << st.committed << "\n"
      << " Rollback efficiency            : "
      << (st.processed ? ((float)(st.processed - st.antis_sent)) / ((float)st.processed) : 1.0f) << "\n"
      << " rollbacks " << st.rollbacks << ", undone " << st.undone << ", annihilated " << st.annihilated
      << ", duplicates " << st.duplicates << ", irregular LP turns " << st.fix_turns << ", sent to other partitions " << st.remote_sent << "\n"
      << " committed digest: " << dg[0] << " " << dg[1] << " " << dg[2] << " " << dg[3] << "\n"
      << "====================================================================\n";
  ssb_destroy(engine_);
  engine_ = NULL;
  for (size_t i = 0; i < comm_files_.size(); ++i) std::remove(comm_files_[i].c_str());      // peers imported them long ago
}

// host handler vs device twin: advance the chains of the first k initial events with App::event_handler on the host
// and look every visited event up among the committed records
template <class App>
void runner<App>::verify_twin(const ev_vec<App>& initial, const std::vector<ssb_record>& committed, long k) {
  if (world_ != 1) { LOG(INFO) << "SCALESIM_B200_VERIFY needs a single partition; skipped"; return; }
  std::map<std::pair<long, long>, const ssb_record*> idx;     // (recv_time, id) -> record
  for (size_t i = 0; i < committed.size(); ++i) idx[std::make_pair((long)committed[i].recv_time, (long)committed[i].id)] = &committed[i];
  st_vec<App> states;
  parti_ptr all(new std::vector<long>());
  std::pair<parti_ptr, parti_indx_ptr> parti = app_.init_partition_index(1);
  app_.init_states_in_this_rank(states, 0, 1, parti.first);
  std::map<long, st_ptr<App> > st_of;
  for (size_t i = 0; i < states.size(); ++i) st_of[states[i]->id()] = states[i];
  long checked = 0, bad = 0;
  for (long c = 0; c < k && c < (long)initial.size(); ++c) {
    ev_ptr<App> e = initial[(size_t)c];
    while (e && e->receive_time() < App::finish_time()) {
      typename std::map<std::pair<long, long>, const ssb_record*>::iterator f =
          idx.find(std::make_pair(e->receive_time(), e->id()));
      ++checked;
      if (f == idx.end() || f->second->dst != e->destination() || f->second->src != e->source() ||
          f->second->send_time != e->send_time()) {
        if (bad++ < 5) LOG(ERROR) << "device twin mismatch at event id " << e->id() << " t " << e->receive_time();
      }
      typename std::map<long, st_ptr<App> >::iterator s = st_of.find(e->destination());
      if (s == st_of.end()) break;
      boost::optional<std::pair<ev_vec<App>, st_ptr<App> > > up = app_.event_handler(e, s->second);
      if (!up || up->first.empty()) break;
      e = up->first[0];
    }
  }
  LOG(INFO) << "host-handler check of the device twin: " << checked << " events, " << bad << " mismatches";
  if (bad) std::exit(3);
}

}  // namespace scalesim
#endif
