// Stand-in for <boost/thread/mutex.hpp>.
#pragma once
#include <mutex>
namespace boost {
using std::mutex;
using std::lock_guard;
using std::unique_lock;
}  // namespace boost
