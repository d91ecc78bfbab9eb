// device_model<phold> — adapter for ScaleSim's PHOLD benchmark application (reference src/phold/phold.hpp).
// Include AFTER the unmodified phold.hpp: it reads that header's file-static parameters and tables
// (NUM_LP, LOOK_AHEAD, RAND_TABLE_SIZE, LATENCY_TABLE, REMOTE_COM_TABLE — phold.hpp:37-58) by name.
// Device twin of phold::event_handler: PholdModel in csrc/ssb_models.cuh.
#ifndef SCALESIM_B200_MODELS_PHOLD_MODEL_HPP_
#define SCALESIM_B200_MODELS_PHOLD_MODEL_HPP_
#include <algorithm>
#include "scalesim/simulation/device_model.hpp"

namespace scalesim {
template <>
struct device_model<phold> {
  struct context {};
  static int model_id() { return SSB_MODEL_PHOLD; }
  static long default_bucket_ticks() { return LOOK_AHEAD; }   // every hop advances time by >= LOOK_AHEAD
  static long n_lp(const parti_ptr& partition) { return (long)partition->size(); }
  static void pack_event(const phold::Event& e, ssb_event& out, context&) {
    out.id = e.id(); out.src = e.source(); out.dst = e.destination();
    out.recv_time = e.receive_time(); out.send_time = e.send_time();
    out.p0 = e.num_hops(); out.p1 = 0; out.flags = e.is_cancel() ? 1 : 0;
  }
  // upper bound of the committed events of a run, for sizing the device log: a chain advances by >= LOOK_AHEAD per hop
  static long committed_bound(const std::vector<ssb_event>& initial, long finish_time) {
    const long hops = LOOK_AHEAD > 0 ? finish_time / LOOK_AHEAD + 1 : 64;
    return (long)initial.size() * std::min<long>(hops, 8);      // ~4.6 hops per chain at lambda 1.0; SCALESIM_B200_MAX_COMMITTED overrides
  }
  static long committed_bound_one(const ssb_raw_record&, long finish_time) {
    return std::min<long>(LOOK_AHEAD > 0 ? finish_time / LOOK_AHEAD + 1 : 64, 8);
  }
  static int state_words() { return 2; }
  static void pack_state(const phold::State& s, uint32_t* w, context&) {
    w[0] = (uint32_t)((uint64_t)s.id() & 0xFFFFFFFFull); w[1] = (uint32_t)((uint64_t)s.id() >> 32);
  }
  // exact-differential store: PHOLD events carry no references into model data
  static std::vector<int64_t> model_data_blob(const context&) { return std::vector<int64_t>(); }
  static void restore_model_data(context&, const std::vector<int64_t>&) {}
  // scenario image (ssb_scenario): the three slots upload_model_data sends
  static void image_slots(context&, ssb_scenario& sc) {
    static int64_t params[2];
    params[0] = LOOK_AHEAD; params[1] = NUM_LP;
    sc.slot_data[0] = LATENCY_TABLE; sc.slot_bytes[0] = (int64_t)(sizeof(long) * (size_t)RAND_TABLE_SIZE);
    sc.slot_data[1] = REMOTE_COM_TABLE; sc.slot_bytes[1] = (int64_t)(sizeof(int) * (size_t)RAND_TABLE_SIZE);
    sc.slot_data[2] = params; sc.slot_bytes[2] = (int64_t)sizeof(params);
  }
  static std::vector<int64_t> model_data_blob_of(const ssb_scenario&) { return std::vector<int64_t>(); }
  static int upload_model_data(ssb_engine* e, context&) {
    static_assert(sizeof(long) == 8 && sizeof(int) == 4, "LP64 expected");
    int rc = ssb_set_model_data(e, 0, LATENCY_TABLE, sizeof(long) * (size_t)RAND_TABLE_SIZE);
    if (rc) return rc;
    rc = ssb_set_model_data(e, 1, REMOTE_COM_TABLE, sizeof(int) * (size_t)RAND_TABLE_SIZE);
    if (rc) return rc;
    int64_t params[2] = {LOOK_AHEAD, NUM_LP};
    return ssb_set_model_data(e, 2, params, sizeof(params));
  }
};
}  // namespace scalesim
#endif
