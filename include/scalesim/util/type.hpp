// Type aliases of the application surface (reference include/scalesim/util/type.hpp:18-31).
#ifndef SCALESIM_B200_UTIL_TYPE_HPP_
#define SCALESIM_B200_UTIL_TYPE_HPP_
#include <vector>
#include <boost/shared_ptr.hpp>
#include <boost/unordered_map.hpp>
#include <boost/unordered_set.hpp>
template <class App> using event = typename App::Event;
template <class App> using state = typename App::State;
template <class App> using ev_ptr = boost::shared_ptr<const event<App> >;
template <class App> using st_ptr = boost::shared_ptr<const state<App> >;
template <class App> using ev_vec = std::vector<ev_ptr<App> >;
template <class App> using st_vec = std::vector<st_ptr<App> >;
using parti_ptr = boost::shared_ptr<std::vector<long> >;         // lp id -> partition
using parti_indx = boost::unordered_multimap<long, long>;         // partition -> lp ids
using parti_indx_ptr = boost::shared_ptr<parti_indx>;
#endif
