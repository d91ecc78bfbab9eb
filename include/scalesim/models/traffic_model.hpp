// device_model<traffic_sim> — adapter for ScaleSim's TrafficSim application (reference src/trafficsim/traffic_sim.hpp).
// Include AFTER the unmodified traffic_sim.hpp. The application keeps the vehicle's remaining route and the
// junction's road list in private vectors; the adapter reads them through the classes' own serialize() members
// (traffic_sim.hpp:84-95, 130-141) with scalesim::field_capture. Routes go to one pool in device memory and events
// carry (offset, nodes left) instead of a vector. Device twin of traffic_sim::event_handler: TrafficModel in
// csrc/ssb_models.cuh.
#ifndef SCALESIM_B200_MODELS_TRAFFIC_MODEL_HPP_
#define SCALESIM_B200_MODELS_TRAFFIC_MODEL_HPP_
#include <stdexcept>
#include "scalesim/simulation/device_model.hpp"

namespace scalesim {
template <>
struct device_model<traffic_sim> {
  static const int kInlineRoads = 2;   // TrafficModel::kInlineRoads
  struct context { std::vector<int64_t> route_pool, road_table; };
  static int model_id() { return SSB_MODEL_TRAFFIC; }
  // shortest possible hop: length >= 1 m, so len*3600/speed/1000 >= 0, plus the fixed 5 s (traffic_sim.hpp:314-316)
  static long default_bucket_ticks() { return 5; }
  static long n_lp(const parti_ptr& partition) { return (long)partition->size(); }
  static void pack_event(const traffic_sim::Event& e, ssb_event& out, context& cx) {
    // serialize order: vehicle_id, arrival_time, departure_time, source_, destinations_, base_{white, cancel}
    field_capture fc = capture_fields(e);
    const std::vector<long>& route = fc.vectors.at(0);
    out.id = e.id(); out.src = e.source(); out.dst = e.destination();
    out.recv_time = e.receive_time(); out.send_time = e.send_time();
    out.p0 = (int64_t)cx.route_pool.size(); out.p1 = (int64_t)route.size();
    out.flags = e.is_cancel() ? 1 : 0;
    cx.route_pool.insert(cx.route_pool.end(), route.begin(), route.end());
  }
  // upper bound of the committed events of a run, for sizing the device log: one event per route node
  static long committed_bound(const std::vector<ssb_event>& initial, long) {
    long n = 0;
    for (size_t i = 0; i < initial.size(); ++i) n += (long)initial[i].p1 + 1;
    return n;
  }
  static long committed_bound_one(const ssb_raw_record& e, long) { return (long)e.p1 + 1; }
  static int state_words() { return 2 + 3 * kInlineRoads; }
  // up to kInlineRoads out-roads sit in the state words; a junction with more keeps them in the road table (model
  // data slot 1) and its state holds the index of its first road there — any junction degree (the reference's State
  // holds unbounded vectors, traffic_sim.hpp:99-142)
  static void pack_state(const traffic_sim::State& s, uint32_t* w, context& cx) {
    // serialize order: id_, destinations_, speed_limit_, num_lanes_, road_length_
    field_capture fc = capture_fields(s);
    const std::vector<long>&dst = fc.vectors.at(0), &speed = fc.vectors.at(1), &len = fc.vectors.at(3);
    for (int i = 0; i < state_words(); ++i) w[i] = 0;
    w[0] = (uint32_t)s.id(); w[1] = (uint32_t)dst.size();
    if (dst.size() <= (size_t)kInlineRoads) {
      for (size_t i = 0; i < dst.size(); ++i) {
        w[2 + 3 * i] = (uint32_t)dst[i]; w[3 + 3 * i] = (uint32_t)speed[i]; w[4 + 3 * i] = (uint32_t)len[i];
      }
    } else {
      w[2] = (uint32_t)(cx.road_table.size() / 3);
      for (size_t i = 0; i < dst.size(); ++i) {
        cx.road_table.push_back(dst[i]); cx.road_table.push_back(speed[i]); cx.road_table.push_back(len[i]);
      }
    }
  }
  // exact-differential store: the recorded events index into the route pool, so it travels with the history
  static std::vector<int64_t> model_data_blob(const context& cx) { return cx.route_pool; }
  static void restore_model_data(context& cx, const std::vector<int64_t>& blob) { cx.route_pool = blob; }
  // scenario image (ssb_scenario): slot 0 = route pool, slot 1 = road table — what upload_model_data sends
  static void image_slots(context& cx, ssb_scenario& sc) {
    if (cx.route_pool.empty()) cx.route_pool.push_back(0);
    sc.slot_data[0] = cx.route_pool.data(); sc.slot_bytes[0] = (int64_t)(cx.route_pool.size() * sizeof(int64_t));
    if (!cx.road_table.empty()) { sc.slot_data[1] = cx.road_table.data(); sc.slot_bytes[1] = (int64_t)(cx.road_table.size() * sizeof(int64_t)); }
  }
  static std::vector<int64_t> model_data_blob_of(const ssb_scenario& sc) {
    const int64_t* p = (const int64_t*)sc.slot_data[0];
    return std::vector<int64_t>(p, p + sc.slot_bytes[0] / 8);
  }
  static int upload_model_data(ssb_engine* e, context& cx) {
    if (cx.route_pool.empty()) cx.route_pool.push_back(0);
    int rc = ssb_set_model_data(e, 0, cx.route_pool.data(), cx.route_pool.size() * sizeof(int64_t));
    if (rc || cx.road_table.empty()) return rc;
    return ssb_set_model_data(e, 1, cx.road_table.data(), cx.road_table.size() * sizeof(int64_t));
  }
};
}  // namespace scalesim
#endif
