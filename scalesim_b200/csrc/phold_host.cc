// phold_host.cc — host-side mirror of the PHOLD application's input generation (no device work).
// Used by callers that drive the C ABI directly (Python tests / bench); the drop-in C++ path gets the
// same data from the unmodified phold.hpp.
//   tables      : src/phold/phold.hpp:144-164 — std::default_random_engine (libstdc++: minstd_rand0),
//                 std::exponential_distribution<double>(lambda) * 100000 truncated to long, and
//                 std::uniform_real_distribution<double>(0,1) < ratio, BOTH seeded with the same seed.
//   init events : src/phold/phold.hpp:197-209 — event id at LP id % NUM_LP, send = receive =
//                 LATENCY_TABLE[id % RAND_TABLE_SIZE], hops 0.
#include <random>
#include "../../include/scalesim_b200.h"

extern "C" {

int ssb_phold_build_tables(int64_t seed, double lambda, double remote_ratio, int64_t n, int64_t* latency, int32_t* remote) {
  if (!latency || !remote || n < 0) return SSB_ERR_INVALID;
  const int kEffectiveDecimal = 100000;   // phold.hpp:54
  {
    std::default_random_engine gen((int)seed);
    std::exponential_distribution<double> dist(lambda);
    for (int64_t i = 0; i < n; ++i) latency[i] = (long)(dist(gen) * kEffectiveDecimal);
  }
  {
    std::default_random_engine gen((int)seed);
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    for (int64_t i = 0; i < n; ++i) remote[i] = dist(gen) < remote_ratio ? 1 : 0;
  }
  return SSB_OK;
}

int64_t ssb_phold_init_events(int64_t num_lp, int64_t num_init_msg, const int64_t* latency, int64_t table_size,
                              int64_t offset, int64_t stride, ssb_event* out, int64_t cap) {
  if (!latency || !out || stride < 1 || num_lp < 1 || table_size < 1) return -1;
  int64_t k = 0;
  for (int64_t id = offset; id < num_init_msg; id += stride) {
    if (k >= cap) return -1;
    ssb_event& e = out[k++];
    e.id = id; e.src = id % num_lp; e.dst = id % num_lp;
    e.recv_time = latency[id % table_size]; e.send_time = e.recv_time;
    e.p0 = 0; e.p1 = 0; e.flags = 0;
  }
  return k;
}

}  // extern "C"
