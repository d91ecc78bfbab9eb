// scalesim/util.hpp — mirror of reference include/scalesim/util.hpp:11-17 (the parts the application surface uses).
#ifndef SCALESIM_B200_UTIL_HPP_
#define SCALESIM_B200_UTIL_HPP_
#include "scalesim/util/flags.hpp"
#include "scalesim/util/timestamp.hpp"
#include "scalesim/util/type.hpp"
#endif
