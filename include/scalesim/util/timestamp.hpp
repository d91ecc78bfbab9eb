// scalesim::timestamp — (time, id) with lexicographic order: time first, then id
// (reference include/scalesim/util/timestamp.hpp:17-83, pinned by test/small/util_test.cc:15-38).
// On the device the same order is the unsigned order of the packed key time << 32 | id.
#ifndef SCALESIM_B200_UTIL_TIMESTAMP_HPP_
#define SCALESIM_B200_UTIL_TIMESTAMP_HPP_
#include <limits>
#include <tuple>
namespace scalesim {
class timestamp {
 public:
  timestamp() : time_(-1), id_(-1) {}
  timestamp(long time, long id) : time_(time), id_(id) {}
  long time() const { return time_; }
  long id() const { return id_; }
  void update(long time, long id) { time_ = time; id_ = id; }
  void update(const timestamp& t) { *this = t; }
  bool operator==(const timestamp& r) const { return time_ == r.time_ && id_ == r.id_; }
  bool operator!=(const timestamp& r) const { return !(*this == r); }
  bool operator<(const timestamp& r) const { return std::tie(time_, id_) < std::tie(r.time_, r.id_); }
  bool operator>(const timestamp& r) const { return r < *this; }
  static timestamp max() { return timestamp(std::numeric_limits<long>::max(), std::numeric_limits<long>::max()); }
  static timestamp zero() { return timestamp(0, 0); }
  static timestamp null() { return timestamp(-1, -1); }
  template <class Archive> void serialize(Archive& ar, unsigned int) { ar & time_; ar & id_; }
 private:
  long time_, id_;
};
}  // namespace scalesim
#endif
