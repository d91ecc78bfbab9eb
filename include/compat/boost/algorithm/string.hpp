// Stand-in for <boost/algorithm/string.hpp>: split() + is_any_of() only, with
// Boost's default token_compress_off behaviour (empty tokens are kept).
#pragma once
#include <algorithm>   // the real header pulls it in; traffic_sim.hpp relies on that for std::find
#include <string>
#include <vector>
namespace boost {
struct is_any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string& s) { return is_any_of_pred{s}; }
template <class Seq, class Pred>
Seq& split(Seq& out, const std::string& in, Pred pred) {
  out.clear();
  std::string cur;
  for (char c : in) {
    if (pred(c)) { out.push_back(cur); cur.clear(); } else { cur.push_back(c); }
  }
  out.push_back(cur);
  return out;
}
namespace algorithm { using boost::split; using boost::is_any_of; }
}  // namespace boost
