// Stand-in for <boost/archive/binary_iarchive.hpp>; reads what the stand-in
// binary_oarchive writes.
#pragma once
#include <cstdint>
#include <istream>
#include <string>
#include <type_traits>
#include <vector>
#include "boost/serialization/access.hpp"
namespace boost { namespace archive {
class binary_iarchive {
 public:
  explicit binary_iarchive(std::istream& is) : is_(is) {}
  typedef std::false_type is_saving;
  typedef std::true_type is_loading;
  template <class T> binary_iarchive& operator>>(T& t) { load(t); return *this; }
  template <class T> binary_iarchive& operator&(T& t) { load(t); return *this; }
 private:
  std::istream& is_;
  template <class T>
  typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value>::type
  load(T& t) { is_.read(reinterpret_cast<char*>(&t), sizeof(T)); }
  template <class T>
  typename std::enable_if<std::is_class<T>::value>::type
  load(T& t) { boost::serialization::access::serialize(*this, t, 0u); }
  template <class T, std::size_t N> void load(T (&a)[N]) { for (std::size_t i = 0; i < N; ++i) load(a[i]); }
  template <class T> void load(std::vector<T>& v) {
    std::uint64_t n = 0; load(n);
    v.resize(n);
    for (auto& x : v) load(x);
  }
  void load(std::vector<bool>& v) {
    std::uint64_t n = 0; load(n);
    v.resize(n);
    for (std::uint64_t i = 0; i < n; ++i) { char c = 0; load(c); v[i] = c != 0; }
  }
  void load(std::string& s) {
    std::uint64_t n = 0; load(n);
    s.resize(n);
    is_.read(&s[0], static_cast<std::streamsize>(n));
  }
};
}}  // namespace boost::archive
