// Command line flags of the reference (include/scalesim/util/flags.hpp:17-21): same names, same defaults.
#ifndef SCALESIM_B200_UTIL_FLAGS_HPP_
#define SCALESIM_B200_UTIL_FLAGS_HPP_
#include <gflags/gflags.h>
DEFINE_bool(diff_init, false, "Initial Execution of Differential Simulation");
DEFINE_bool(diff_repeat, false, "Repeat Execution of Differential Simulation");
DEFINE_string(sim_id, "0", "Execution ID for Differential Simulation");
DEFINE_string(store_dir, "", "Directory PATH for Event/State Stores");
DEFINE_string(store_ip, "", "IP address for Event/State Stores");
#endif
