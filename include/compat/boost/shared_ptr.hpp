// Stand-in for <boost/shared_ptr.hpp>: std::shared_ptr under the boost:: name.
// Part of scalesim-b200's compat include dir: ScaleSim application sources
// (phold.hpp, traffic_sim.hpp) name Boost directly; Boost is not a dependency here.
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::static_pointer_cast;
using std::const_pointer_cast;
}  // namespace boost
