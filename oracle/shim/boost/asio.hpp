// Reference-harness stand-in for <boost/asio.hpp>: just enough io_service
// (post / run / stop + the `work` keep-alive) for scalesim::thr_pool
// (reference include/scalesim/util/thread_manager.hpp:19-60).
// TEST INFRASTRUCTURE ONLY: used to build the unmodified reference into oracle/_ref.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <deque>
#include <functional>
#include <mutex>
namespace boost { namespace asio {
class io_service {
 public:
  class work {
   public:
    explicit work(io_service& s) : s_(s) { std::lock_guard<std::mutex> g(s_.m_); ++s_.work_; }
    ~work() { { std::lock_guard<std::mutex> g(s_.m_); --s_.work_; } s_.cv_.notify_all(); }
   private:
    io_service& s_;
  };
  template <class F> void post(F f) {
    { std::lock_guard<std::mutex> g(m_); q_.emplace_back(f); }
    cv_.notify_one();
  }
  std::size_t run() {
    std::size_t n = 0;
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stopped_ || !q_.empty() || work_ == 0; });
        if (stopped_) return n;
        if (q_.empty()) { if (work_ == 0) return n; continue; }
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
      ++n;
    }
  }
  void stop() { { std::lock_guard<std::mutex> g(m_); stopped_ = true; } cv_.notify_all(); }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()> > q_;
  int work_ = 0;
  bool stopped_ = false;
};
}}  // namespace boost::asio
