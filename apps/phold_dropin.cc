// Drop-in build of ScaleSim's PHOLD: the reference's application header and its main() are compiled UNMODIFIED
// (read in place, nothing copied); only the include path differs — "scalesim/simulation.hpp" resolves to this
// repo's host mirror, and the device adapter is added between the application header and main().
#include "scalesim/simulation.hpp"            // this repo (include/scalesim/...)
#include "phold/phold.hpp"                    // reference src/phold/phold.hpp, unmodified
#include "scalesim/models/phold_model.hpp"    // device_model<phold>
#include "phold/main_phold.cc"                // reference src/phold/main_phold.cc, unmodified (its includes are guarded)
