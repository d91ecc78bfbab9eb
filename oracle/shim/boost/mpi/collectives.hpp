#pragma once
#include "boost/mpi.hpp"
