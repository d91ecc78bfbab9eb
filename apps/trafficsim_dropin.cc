// Drop-in build of ScaleSim's TrafficSim: reference sources compiled unmodified, see phold_dropin.cc.
#include "scalesim/simulation.hpp"                 // this repo
#include "trafficsim/traffic_sim.hpp"              // reference src/trafficsim/traffic_sim.hpp, unmodified
#include "scalesim/models/traffic_model.hpp"       // device_model<traffic_sim>
#include "trafficsim/main_trafficsim.cc"           // reference src/trafficsim/main_trafficsim.cc, unmodified
