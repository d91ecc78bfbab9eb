// Reference-harness stand-in for <boost/mpi.hpp>: a ONE-RANK world. Sends to self
// are queued and handed back by iprobe/recv; collectives are identities.
// TEST INFRASTRUCTURE ONLY: used to build the unmodified reference into oracle/_ref
// (the reference's singletons allow one rank per process, so this is a single-rank
// harness: `mpirun -np 1`).
#pragma once
#include <algorithm>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>
namespace boost { namespace mpi {
enum { any_source = -1, any_tag = -1 };
class environment {
 public:
  environment() {}
  environment(int&, char**&) {}
};
class request {
 public:
  bool done = true;
};
class communicator {
 public:
  int rank() const { return 0; }
  int size() const { return 1; }
  void barrier() const {}
  template <class T> request isend(int, int, const T& v) const {
    std::lock_guard<std::mutex> g(box().m);
    box().q.emplace_back(std::make_shared<holder<T> >(v));
    return request();
  }
  bool iprobe(int = any_source, int = any_tag) const {
    std::lock_guard<std::mutex> g(box().m);
    return !box().q.empty();
  }
  template <class T> void recv(int, int, T& out) const {
    std::shared_ptr<base> b;
    { std::lock_guard<std::mutex> g(box().m); b = box().q.front(); box().q.pop_front(); }
    out = static_cast<holder<T>*>(b.get())->v;
  }
 private:
  struct base { virtual ~base() {} };
  template <class T> struct holder : base { explicit holder(const T& x) : v(x) {} T v; };
  struct mailbox { std::mutex m; std::deque<std::shared_ptr<base> > q; };
  static mailbox& box() { static mailbox b; return b; }
};
template <class It> It test_some(It first, It last) { (void)last; return first; }
template <class T> struct minimum { const T& operator()(const T& a, const T& b) const { return b < a ? b : a; } };
template <class T> struct maximum { const T& operator()(const T& a, const T& b) const { return a < b ? b : a; } };
template <class T, class Op> T all_reduce(const communicator&, const T& v, Op) { return v; }
template <class T>
void all_to_all(const communicator&, const std::vector<T>& in, int n, std::vector<T>& out) {
  out.assign(in.begin(), in.begin() + n);
}
}}  // namespace boost::mpi
