// Stand-in for <boost/filesystem.hpp>.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem = std::filesystem; }
