// Stand-in for <boost/archive/binary_oarchive.hpp>: a self-consistent little-endian
// byte format (NOT Boost's wire format; the reference pins no archive version and no
// test inspects value bytes). Handles arithmetic types, std::vector, std::string,
// C arrays and classes exposing serialize(Archive&, unsigned).
#pragma once
#include <cstdint>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>
#include "boost/serialization/access.hpp"
namespace boost { namespace archive {
class binary_oarchive {
 public:
  explicit binary_oarchive(std::ostream& os) : os_(os) {}
  typedef std::true_type is_saving;
  typedef std::false_type is_loading;
  template <class T> binary_oarchive& operator<<(const T& t) { save(t); return *this; }
  template <class T> binary_oarchive& operator&(const T& t) { save(t); return *this; }
 private:
  std::ostream& os_;
  template <class T>
  typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value>::type
  save(const T& t) { os_.write(reinterpret_cast<const char*>(&t), sizeof(T)); }
  template <class T>
  typename std::enable_if<std::is_class<T>::value>::type
  save(const T& t) { boost::serialization::access::serialize(*this, const_cast<T&>(t), 0u); }
  template <class T, std::size_t N> void save(const T (&a)[N]) { for (std::size_t i = 0; i < N; ++i) save(a[i]); }
  template <class T> void save(const std::vector<T>& v) {
    std::uint64_t n = v.size(); save(n);
    for (const auto& x : v) save(x);
  }
  void save(const std::vector<bool>& v) {
    std::uint64_t n = v.size(); save(n);
    for (bool b : v) { char c = b; save(c); }
  }
  void save(const std::string& s) {
    std::uint64_t n = s.size(); save(n);
    os_.write(s.data(), static_cast<std::streamsize>(n));
  }
};
}}  // namespace boost::archive
