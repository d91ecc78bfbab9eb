// Stand-in for <boost/make_shared.hpp>.
#pragma once
#include "boost/shared_ptr.hpp"
