// Stand-in for <boost/optional.hpp>: std::optional under the boost:: name.
#pragma once
#include <optional>
namespace boost {
template <class T> using optional = std::optional<T>;
}  // namespace boost
