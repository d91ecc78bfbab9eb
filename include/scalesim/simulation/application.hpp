// scalesim::application — base class of an application with its static tunables
// (reference include/scalesim/simulation/application.hpp:18-56). Only finish_time() changes results; the scheduling
// knobs of the CPU runtime (gsync_interval, switch_lp_interval, global_cut_interval, num_thr) are accepted and ignored:
// the device engine works in windows (see ssb_config.bucket_ticks / window_buckets).
#ifndef SCALESIM_B200_SIMULATION_APPLICATION_HPP_
#define SCALESIM_B200_SIMULATION_APPLICATION_HPP_
#include <limits>
#include <boost/thread.hpp>
#include "scalesim/util.hpp"
namespace scalesim {
class application {
 private:
  application(const application&);
  void operator=(const application&);
 protected:
  application() {}
  virtual ~application() {}
 public:
  static int gsync_interval() { return 100; }
  static int switch_lp_interval() { return 5; }
  static int global_cut_interval() { return 20; }
  static int num_thr() { return boost::thread::physical_concurrency() - 1; }
  static long finish_time() { return std::numeric_limits<long>::max(); }
};
}  // namespace scalesim
#endif
