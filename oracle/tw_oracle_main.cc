/*
 * tw_oracle_main.cc — command line front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 *   tw_oracle phold <NUM_LP> <NUM_INIT_MSG> <REMOTE_RATIO> <LAMBDA> [options]
 *   tw_oracle traffic <road csv> <trip csv> [options]
 * options: --timewarp <num_sched>  sequential Time Warp emulation (default: 4 schedulers)
 *          --closure <threads>     closure of the initial events' chains
 *          --print                 write "RE,..." lines to stdout (reference format, sim_obj.hpp:66-77)
 * stderr gets: committed count, digest, stats, elapsed seconds.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "tw_oracle.h"

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: see tw_oracle_main.cc\n"); return 2; }
  two_sim* s = nullptr;
  int pos = 1;
  std::string app = argv[pos++];
  if (app == "phold") {
    if (argc < 6) return 2;
    long n = atol(argv[pos++]); long m = atol(argv[pos++]);
    double ratio = atof(argv[pos++]); double lambda = atof(argv[pos++]);
    s = two_sim_phold(n, m, ratio, lambda, 1);
  } else if (app == "traffic") {
    const char* rd = argv[pos++]; const char* trip = argv[pos++];
    s = two_sim_traffic_files(rd, trip);
    if (!s) { std::fprintf(stderr, "cannot open inputs\n"); return 1; }
  } else return 2;
  int mode_tw = 4, closure_thr = 0; bool print = false;
  for (; pos < argc; ++pos) {
    if (!strcmp(argv[pos], "--timewarp") && pos + 1 < argc) { mode_tw = atoi(argv[++pos]); closure_thr = 0; }
    else if (!strcmp(argv[pos], "--closure") && pos + 1 < argc) { closure_thr = atoi(argv[++pos]); }
    else if (!strcmp(argv[pos], "--print")) print = true;
  }
  auto t0 = std::chrono::steady_clock::now();
  long n;
  if (closure_thr > 0) n = two_sim_run_closure(s, closure_thr, print ? 1 : 0);
  else n = two_sim_run_timewarp(s, mode_tw, 100, app == "phold" ? 1 : 5);
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t d[4]; int64_t st[4];
  two_sim_digest(s, d); two_sim_stats(s, st);
  std::fprintf(stderr, "committed %ld digest %016lx %016lx %016lx processed %ld cancels %ld gvt_rounds %ld remote %ld elapsed %.3f s\n",
               n, d[1], d[2], d[3], (long)st[0], (long)st[1], (long)st[2], (long)st[3], sec);
  if (print) {
    std::vector<two_record> r((size_t)two_sim_num_committed(s));
    long k = two_sim_committed(s, r.data(), (long)r.size());
    for (long i = 0; i < k; ++i) {
      std::printf("RE,%ld,%ld,%ld,%ld,%ld%s\n", (long)r[i].id, (long)r[i].src, (long)r[i].send, (long)r[i].dst,
                  (long)r[i].recv, r[i].tag == 1 ? ",START" : (r[i].tag == 2 ? ",END" : ""));
    }
  }
  two_sim_free(s);
  return 0;
}
