// Stand-in for <boost/serialization/access.hpp>: the friend hook that lets an
// archive call a class's private serialize(Archive&, unsigned).
#pragma once
namespace boost { namespace serialization {
class access {
 public:
  template <class Archive, class T>
  static void serialize(Archive& ar, T& t, unsigned int version) { t.serialize(ar, version); }
};
}}  // namespace boost::serialization
