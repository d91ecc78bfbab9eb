// scalesim::runner<App> — same entry points as the reference's driver (include/scalesim/simulation/runner.hpp:28-60:
// init_main(&argc, &argv), run()), but the Time Warp core behind it is the device engine of libscalesim_b200.so:
//
//   run():  App::init()                                      (runner.hpp:82)
//           App::init_partition_index(world)  -> ssb_set_partition            (runner.hpp:98)
//           App::init_states_in_this_rank     -> ssb_load_states              (runner.hpp:108-118)
//           App::init_events + shuffle by partition[destination] -> ssb_load_events   (runner.hpp:131-147)
//           ssb_run                           = runner::loop                   (runner.hpp:350-396)
//           ssb_drain_committed -> "RE,..." lines on stdout                     (queue.hpp:203-211, sim_obj.hpp:66-77)
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
#ifndef SCALESIM_B200_SIMULATION_RUNNER_HPP_
#define SCALESIM_B200_SIMULATION_RUNNER_HPP_

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <string>
#include <thread>
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
  // exact-differential history store of this rank: <store_dir>/<sim_id>/history<rank>.ssb (same file format as
  // scalesim_b200/store.py): the committed events with their sent-log entries, the LPs that had an initial state and
  // the model data the events refer to (TrafficSim: the route pool)
  struct history {
    std::vector<ssb_raw_record> rec;
    std::vector<ssb_sent_record> sent;
    std::vector<int64_t> state_lps, pool;
  };
  std::string history_path() const {
    return (FLAGS_store_dir.empty() ? std::string(".") : FLAGS_store_dir) + "/" + FLAGS_sim_id + "/history" + std::to_string(rank_) + ".ssb";
  }
  void write_history(const history& h);
  void read_history(history& h);
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
  int rc = 0;
  if (!p2p) {      // events travel by ncclSend/ncclRecv, GVT by ncclAllReduce: the ranks join one NCCL communicator
  char id[128];
  if (rank_ == 0) {
    rc = ssb_comm_unique_id(id);
    if (rc) fail("ssb_comm_unique_id", rc);
    std::string tmp = std::string(path) + ".tmp";
    { std::ofstream f(tmp.c_str(), std::ios::binary); f.write(id, 128); }
    std::rename(tmp.c_str(), path);
  } else {
    for (int tries = 0;; ++tries) {
      std::ifstream f(path, std::ios::binary);
      if (f && f.read(id, 128) && f.gcount() == 128) break;
      if (tries > 600) { LOG(ERROR) << "timed out waiting for " << path; std::exit(1); }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
  }
  rc = ssb_comm_init(engine_, id);
  if (rc) fail("ssb_comm_init", rc);
  return;
  }
  // peer-memory exchange (default; events AND the GVT reduction cross through the mailboxes, no NCCL): swap the CUDA IPC handles of the engines' mailboxes through files next to the id file
  char mine[64];
  if ((rc = ssb_comm_export(engine_, mine))) fail("ssb_comm_export", rc);
  {
    std::string f = std::string(path) + ".h" + std::to_string(rank_), tmp = f + ".tmp";
    { std::ofstream o(tmp.c_str(), std::ios::binary); o.write(mine, 64); }
    std::rename(tmp.c_str(), f.c_str());
  }
  std::vector<char> all((size_t)world_ * 64);
  for (int r = 0; r < world_; ++r) {
    std::string f = std::string(path) + ".h" + std::to_string(r);
    for (int tries = 0;; ++tries) {
      std::ifstream in(f.c_str(), std::ios::binary);
      if (in && in.read(&all[(size_t)r * 64], 64) && in.gcount() == 64) break;
      if (tries > 600) { LOG(ERROR) << "timed out waiting for " << f; std::exit(1); }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
  }
  if ((rc = ssb_comm_import(engine_, all.data(), world_))) fail("ssb_comm_import", rc);
}

template <class App>
void runner<App>::write_history(const history& h) {
  const std::string p = history_path(), dir = p.substr(0, p.find_last_of('/'));
  std::error_code ec;
  std::filesystem::create_directories(dir, ec);
  if (ec) { LOG(ERROR) << "cannot create " << dir << ": " << ec.message(); std::exit(1); }
  std::ofstream f((p + ".tmp").c_str(), std::ios::binary);
  const int64_t hdr[3] = {(int64_t)h.rec.size(), (int64_t)h.state_lps.size(), (int64_t)h.pool.size()};
  f.write("SSBHIST2", 8);
  f.write((const char*)hdr, sizeof(hdr));
  f.write((const char*)h.rec.data(), (std::streamsize)(h.rec.size() * sizeof(ssb_raw_record)));
  f.write((const char*)h.sent.data(), (std::streamsize)(h.sent.size() * sizeof(ssb_sent_record)));
  f.write((const char*)h.state_lps.data(), (std::streamsize)(h.state_lps.size() * 8));
  f.write((const char*)h.pool.data(), (std::streamsize)(h.pool.size() * 8));
  f.close();
  if (!f || std::rename((p + ".tmp").c_str(), p.c_str()) != 0) { LOG(ERROR) << "cannot write " << p; std::exit(1); }
  LOG(INFO) << " Stored objects size : " << (h.rec.size() * (sizeof(ssb_raw_record) + sizeof(ssb_sent_record)) + h.state_lps.size() * 8)
            << " byte (" << h.rec.size() << " events with their sent-log entries, " << h.state_lps.size() << " initial states) -> " << p;
}

template <class App>
void runner<App>::read_history(history& h) {
  const std::string p = history_path();
  std::ifstream f(p.c_str(), std::ios::binary);
  char magic[8];
  int64_t hdr[3];
  if (!f || !f.read(magic, 8) || std::memcmp(magic, "SSBHIST2", 8) != 0 || !f.read((char*)hdr, sizeof(hdr))) {
    LOG(ERROR) << "no history store at " << p << " (run with --diff_init first)";
    std::exit(1);
  }
  h.rec.resize((size_t)hdr[0]); h.sent.resize((size_t)hdr[0]); h.state_lps.resize((size_t)hdr[1]); h.pool.resize((size_t)hdr[2]);
  f.read((char*)h.rec.data(), (std::streamsize)(h.rec.size() * sizeof(ssb_raw_record)));
  f.read((char*)h.sent.data(), (std::streamsize)(h.sent.size() * sizeof(ssb_sent_record)));
  f.read((char*)h.state_lps.data(), (std::streamsize)(h.state_lps.size() * 8));
  f.read((char*)h.pool.data(), (std::streamsize)(h.pool.size() * 8));
  if (!f) { LOG(ERROR) << p << " is truncated"; std::exit(1); }
}

template <class App>
void runner<App>::run() {
  if (FLAGS_diff_init && FLAGS_diff_repeat) { LOG(ERROR) << "--diff_init and --diff_repeat exclude each other"; std::exit(2); }
  typedef std::chrono::steady_clock clk;
  clk::time_point t0 = clk::now();
  app_.init();
  LOG_IF(INFO, rank_ == 0) << "Initiate application";

  std::pair<parti_ptr, parti_indx_ptr> parti = app_.init_partition_index(world_);
  const std::vector<long>& part = *parti.first;
  const long n_lp = model::n_lp(parti.first);
  typename model::context cx;

  // states of this rank
  st_vec<App> states;
  app_.init_states_in_this_rank(states, rank_, world_, parti.first);
  // events: every rank reads the whole scenario and keeps what it owns (the reference shuffles instead, runner.hpp:131-147)
  ev_vec<App> mine;
  history hist;
  std::vector<ssb_event> deletes;      // anti-messages of the DE / RE what-if queries
  if (!FLAGS_diff_repeat) {
    for (int r = 0; r < world_; ++r) {
      ev_vec<App> parsed;
      app_.init_events(parsed, r, world_);
      for (typename ev_vec<App>::iterator it = parsed.begin(); it != parsed.end(); ++it) {
        long d = (*it)->destination();
        if (d >= 0 && d < n_lp && part[d] % world_ == rank_) mine.push_back(*it);
      }
    }
  } else {
    // --diff_repeat (runner::init_repeat, runner.hpp:178-348): no events are read from the scenario. The recorded
    // history of this rank comes back from the store, the what-if queries (App::init_what_if) are injected on top.
    read_history(hist);
    model::restore_model_data(cx, hist.pool);
    for (int r = 0; r < world_; ++r) {
      std::vector<boost::shared_ptr<const what_if<App> > > wi;
      app_.init_what_if(wi, r, world_);
      for (size_t i = 0; i < wi.size(); ++i) {
        if (wi[i]->update_state_.id() != -1) {
          LOG(ERROR) << "what-if queries of type SC (state change) are not part of this build yet";
          std::exit(2);
        }
        long d = wi[i]->lp_id_;
        if (!(d >= 0 && d < n_lp && part[d] % world_ == rank_)) continue;
        if (wi[i]->delete_event_id_ != -1) {
          // delete event (runner.hpp:246-276): the recorded event loses its sent-log entry — eventq::delete_can: its
          // output is NOT cancelled — and an anti-message goes into the LP's inbox (eventq::delete_ev,
          // queue.hpp:227-241), which erases the event and rolls the LP back to its key
          const uint64_t key = ((uint64_t)wi[i]->time_ << 32) | (uint64_t)(uint32_t)wi[i]->delete_event_id_;
          for (size_t k = 0; k < hist.rec.size(); ++k)
            if (hist.rec[k].key == key && (long)hist.rec[k].dst == d) hist.sent[k].dst = 0xffffffffu;
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
  const size_t n_resident = pod.size() + hist.rec.size();

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
  cfg.max_committed = quiet ? 0 : (1 << 24);
  if (FLAGS_diff_init) cfg.flags |= SSB_FLAG_HISTORY;
  cfg.max_snapshots = 1 << 20;
  cfg.rounds_per_sync = 4;
  int rc = ssb_create(&cfg, &engine_);
  if (rc) { LOG(ERROR) << "ssb_create failed (status " << rc << "): " << ssb_last_error(NULL); std::exit(1); }
  comm_init();
  std::vector<int64_t> part64(part.begin(), part.end());
  for (size_t i = 0; i < part64.size(); ++i) part64[i] %= world_;
  if ((rc = ssb_set_partition(engine_, part64.data(), (int64_t)part64.size()))) fail("ssb_set_partition", rc);
  {
    const int sw = model::state_words();
    std::vector<int64_t> ids(states.size());
    std::vector<uint32_t> words(states.size() * (size_t)sw);
    for (size_t i = 0; i < states.size(); ++i) {
      ids[i] = states[i]->id();
      model::pack_state(*states[i], &words[i * (size_t)sw], cx);      // may add to the model data (TrafficSim road table)
    }
    if ((rc = model::upload_model_data(engine_, cx))) fail("ssb_set_model_data", rc);
    if ((rc = ssb_load_states(engine_, ids.data(), words.data(), (int64_t)ids.size()))) fail("ssb_load_states", rc);
  }
  if (FLAGS_diff_repeat &&
      (rc = ssb_load_history(engine_, hist.rec.data(), hist.sent.data(), (int64_t)hist.rec.size()))) fail("ssb_load_history", rc);
  if ((rc = ssb_load_events(engine_, pod.data(), (int64_t)pod.size()))) fail("ssb_load_events", rc);
  if (FLAGS_diff_init) {
    for (size_t i = 0; i < states.size(); ++i) hist.state_lps.push_back(states[i]->id());
  }
  double init_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  LOG_IF(INFO, rank_ == 0)
      << "\n====================================================================\n"
      << " Number of partitions: " << world_ << " (one B200 each)\n"
      << "     Total events num (this partition): " << pod.size() << "\n"
      << "     Total states num (this partition): " << states.size() << "\n"
      << " Total init time: " << (long)init_ms << " ms\n"
      << "====================================================================\n";

  LOG_IF(INFO, rank_ == 0) << "Start Simulation";
  // the committed log is drained while the simulation advances: a window of rounds, then the records so far
  std::vector<ssb_record> buf(1 << 22);
  std::vector<ssb_record> kept;
  const long verify_k = env_long("SCALESIM_B200_VERIFY", 0);
  clk::time_point t1 = clk::now();
  ssb_stats st;
  std::memset(&st, 0, sizeof(st));
  std::vector<char> line(1 << 16);
  std::setvbuf(stdout, NULL, _IOFBF, 1 << 22);
  for (;;) {
    int done = 0;
    for (int i = 0; i < 4 && !done; ++i)
      if ((rc = ssb_step(engine_, &done))) fail("ssb_step", rc);
    if (FLAGS_diff_init) {
      // --diff_init: the committed events leave the device together with their sent-log entries; the lines are printed
      // from the same records and the history accumulates on the host (written to the store after the loop)
      for (;;) {
        const size_t at = hist.rec.size(), cap = 1 << 20;
        hist.rec.resize(at + cap); hist.sent.resize(at + cap);
        int64_t n = 0;
        if ((rc = ssb_drain_history(engine_, &hist.rec[at], &hist.sent[at], (int64_t)cap, &n))) fail("ssb_drain_history", rc);
        hist.rec.resize(at + (size_t)n); hist.sent.resize(at + (size_t)n);
        for (size_t i = at; i < hist.rec.size(); ++i) {
          const ssb_raw_record& r = hist.rec[i];
          std::printf("RE,%ld,%ld,%ld,%ld,%ld%s\n", (long)(r.key & 0xffffffffu), r.src == 0xffffffffu ? -1L : (long)r.src,
                      (long)r.send_time, (long)r.dst, (long)(r.key >> 32), r.tag == 1 ? ",START" : (r.tag == 2 ? ",END" : ""));
        }
        if (n < (int64_t)cap) break;
      }
    } else if (!quiet) {
      for (;;) {
        int64_t n = 0;
        if ((rc = ssb_drain_committed(engine_, buf.data(), (int64_t)buf.size(), &n))) fail("ssb_drain_committed", rc);
        for (int64_t i = 0; i < n; ++i) {
          const ssb_record& r = buf[(size_t)i];
          std::printf("RE,%ld,%ld,%ld,%ld,%ld%s\n", (long)r.id, (long)r.src, (long)r.send_time, (long)r.dst,
                      (long)r.recv_time, r.tag == 1 ? ",START" : (r.tag == 2 ? ",END" : ""));
        }
        if (verify_k > 0) kept.insert(kept.end(), buf.begin(), buf.begin() + n);
        if (n < (int64_t)buf.size()) break;
      }
    }
    if (done) break;
  }
  std::fflush(stdout);
  double sim_ms = std::chrono::duration<double, std::milli>(clk::now() - t1).count();
  if ((rc = ssb_get_stats(engine_, &st))) fail("ssb_get_stats", rc);
  uint64_t dg[4];
  ssb_committed_digest(engine_, dg);
  if (verify_k > 0 && !quiet && !FLAGS_diff_init && !FLAGS_diff_repeat) verify_twin(mine, kept, verify_k);
  if (FLAGS_diff_init) {
    hist.pool = model::model_data_blob(cx);
    write_history(hist);
  }

  LOG(INFO) << "Finish Simulation";
  LOG(INFO)
      << "\n====================================================================\n"
      << " Partition " << rank_ << " of " << world_ << "\n"
      << " Simulation elapsed time: " << (long)sim_ms << " ms (window loop + draining " << (quiet ? "nothing" : "the result lines") << ")\n"
      << "\n"
      << " # of global synchronizations   : " << st.rounds << "\n"
      << " # of processing events         : " << st.processed << "\n"
      << " # of cancels (= # of rollback) : " << st.antis_sent << "\n"
      << " # of output events             : " << st.committed << "\n"
      << " Rollback efficiency            : "
      << (st.processed ? ((float)(st.processed - st.antis_sent)) / ((float)st.processed) : 1.0f) << "\n"
      << " rollbacks " << st.rollbacks << ", undone " << st.undone << ", annihilated " << st.annihilated
      << ", duplicates " << st.duplicates << ", sent to other partitions " << st.remote_sent << "\n"
      << " committed digest: " << dg[0] << " " << dg[1] << " " << dg[2] << " " << dg[3] << "\n"
      << "====================================================================\n";
  ssb_destroy(engine_);
  engine_ = NULL;
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
