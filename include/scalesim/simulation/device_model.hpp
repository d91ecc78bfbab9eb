// device_model<App> — the per-application adapter between ScaleSim's host-side App surface and the device engine.
//
// The reference runs App::event_handler on host objects (boost::shared_ptr<const Event>, std::vector, file-static
// tables); the device engine runs a device twin of that handler (csrc/ssb_models.cuh) on fixed-width records. An
// adapter says which twin to use and how to turn the App's events / states / tables into the C ABI's PODs:
//
//   template <> struct scalesim::device_model<App> {
//     static int model_id();                                     // ssb_model_id
//     static long default_bucket_ticks();                         // window granularity (the app's lookahead)
//     static void pack_event(const App::Event&, ssb_event&, context&);
//     static int state_words();  static void pack_state(const App::State&, uint32_t* words, context&);
//     static int upload_model_data(ssb_engine*, context&);        // ssb_set_model_data calls
//     static long committed_bound(const std::vector<ssb_event>&, long finish_time);   // sizes the device log
//     static long committed_bound_one(const ssb_raw_record&, long finish_time);       // ... for a scenario image's events
//     static void image_slots(context&, ssb_scenario&);           // model data slots of a scenario image to write
//     static std::vector<int64_t> model_data_blob_of(const ssb_scenario&);   // what the history store keeps of them
//   };
//
// Adapters for the two shipped applications live in scalesim/models/; a translation unit that wants the drop-in
// build includes the (unmodified) application header first and the adapter after it (see apps/*.cc).
#ifndef SCALESIM_B200_SIMULATION_DEVICE_MODEL_HPP_
#define SCALESIM_B200_SIMULATION_DEVICE_MODEL_HPP_
#include <cstdint>
#include <type_traits>
#include <vector>
#include <boost/serialization/access.hpp>
#include "scalesim_b200.h"

namespace scalesim {

template <class App> struct device_model;   // specialised per application

// Reads the fields of an object in the order of its (possibly private) serialize(Archive&, unsigned): integers are
// appended to `scalars`, vectors of integers to `vectors`. This is how an adapter reaches fields the application
// class does not expose through accessors (traffic_sim::Event::destinations_, traffic_sim::State::*), without
// touching the application source.
class field_capture {
 public:
  std::vector<long> scalars;
  std::vector<std::vector<long> > vectors;
  typedef std::true_type is_saving;
  typedef std::false_type is_loading;
  template <class T> field_capture& operator&(const T& t) { put(t); return *this; }
  template <class T> field_capture& operator<<(const T& t) { put(t); return *this; }
 private:
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type put(const T& t) { scalars.push_back((long)t); }
  template <class T> typename std::enable_if<std::is_class<T>::value>::type put(const T& t) {
    boost::serialization::access::serialize(*this, const_cast<T&>(t), 0u);
  }
  template <class T> void put(const std::vector<T>& v) { vectors.emplace_back(v.begin(), v.end()); }
};

template <class T> field_capture capture_fields(const T& obj) {
  field_capture fc;
  boost::serialization::access::serialize(fc, const_cast<T&>(obj), 0u);
  return fc;
}

}  // namespace scalesim
#endif
