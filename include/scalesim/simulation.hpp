// scalesim/simulation.hpp — host-side mirror of ScaleSim's application template surface
// (reference include/scalesim/simulation.hpp:11-13). Same header name and same public names, so the
// reference's application sources (src/phold/phold.hpp, src/trafficsim/traffic_sim.hpp, both main_*.cc)
// compile unmodified; behind runner<App> sits libscalesim_b200.so (include/scalesim_b200.h) instead of the
// MPI / boost-thread runtime.
#ifndef SCALESIM_B200_SIMULATION_HPP_
#define SCALESIM_B200_SIMULATION_HPP_
#include "scalesim/util.hpp"
#include "scalesim/simulation/sim_obj.hpp"
#include "scalesim/simulation/application.hpp"
#include "scalesim/simulation/device_model.hpp"
#include "scalesim/simulation/runner.hpp"
#endif
