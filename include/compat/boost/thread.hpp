// Stand-in for <boost/thread.hpp>: std::thread plus the two Boost extras ScaleSim
// uses (thread::physical_concurrency, thread_group). The worker count can be pinned
// with the SCALESIM_NUM_THR environment variable (the real Boost call counts
// physical cores).
#pragma once
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>
#include "boost/shared_ptr.hpp"
#include "boost/thread/mutex.hpp"
namespace boost {
class thread : public std::thread {
 public:
  using std::thread::thread;
  static unsigned physical_concurrency() {
    if (const char* e = std::getenv("SCALESIM_NUM_THR")) {
      int v = std::atoi(e);
      if (v > 0) return static_cast<unsigned>(v);
    }
    unsigned n = std::thread::hardware_concurrency();
    return n ? n : 1;
  }
};
class thread_group {
 public:
  thread_group() {}
  thread_group(const thread_group&) = delete;
  ~thread_group() { join_all(); }
  template <class F> thread* create_thread(F f) {
    threads_.emplace_back(new thread(f));
    return threads_.back().get();
  }
  void join_all() {
    for (auto& t : threads_) if (t->joinable()) t->join();
  }
 private:
  std::vector<std::unique_ptr<thread> > threads_;
};
}  // namespace boost
