// cuda_emu.h — DEVELOPER TOOL, never shipped and never loadable by the product (scalesim_b200/capi.py has no switch
// for it): a sequential host emulation of the CUDA constructs the kernels in scalesim_b200/csrc use, so that their
// LOGIC can be exercised with gdb / AddressSanitizer on a machine without a GPU (tools/emu/build.sh compiles
// ssb_engine.cu with g++ -DSSB_EMU into tools/emu/libssb_emu.so). Every kernel runs as ONE block of ONE thread ("warp"
// of one lane): grid-stride loops visit every element, atomics are plain read-modify-writes, barriers are no-ops.
// The CUDA graph of the window loop is interpreted (ssb_engine::run_graph_emulated). Several partitions = several
// engines of one process, one host thread each (tools/emu/multi.py): shared memory is thread_local, a mailbox's "IPC
// handle" is its address, system fences are real fences, and the six NCCL calls of the NCCL exchange mode have an
// in-process stand-in (emu_nccl below).
// It finds logic errors and out-of-bounds accesses, not races; parity claims are made on the B200 only.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
// (thread_local: several emulated engines — the partitions of a multi-GPU run — may execute kernels at the same time,
// each on its own host thread: tools/emu/multi.py)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct uint2 { unsigned int x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
struct dim3 { unsigned int x = 1, y = 1, z = 1; };
inline dim3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};

template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicXor(T* p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> inline T atomicAdd_system(T* p, T v) { return atomicAdd(p, v); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_sync(unsigned, T v, int) { return v; }
inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
inline bool __any_sync(unsigned, bool p) { return p; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline void __syncthreads() {}
inline void __syncwarp() {}
inline void __threadfence() {}
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }      // a spin on a peer partition's store: let its thread run

// conditional handles of the interpreted graph (ssb_engine::build_graph / run_graph_emulated): small integers indexing
// a per-thread table (one engine per host thread)
typedef unsigned long long cudaGraphConditionalHandle;
inline thread_local unsigned g_emu_cond[4] = {0, 0, 0, 0};
inline void cudaGraphSetConditional(cudaGraphConditionalHandle h, unsigned value) { g_emu_cond[h & 3u] = value; }

// ---- the slice of the runtime API that ssb_engine.cu uses: host memory, no streams ----
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "emulated runtime error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
typedef void* cudaStream_t;
struct cudaEmuEvent { std::chrono::steady_clock::time_point t; };
typedef cudaEmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
constexpr unsigned cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2;
struct cudaDeviceProp { int multiProcessorCount; };
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 1; return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cudaEmuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
// Set when an emulated engine stops with a device error: its peer partitions (other engines of the process, waiting
// for it at the GVT exchange) give up at once instead of spinning out their bound (k_gvt's SSB_EMU branch).
inline std::atomic<int> g_emu_peer_failed{0};
// "IPC" inside one process: the handle of a mailbox is its address (the partitions of an emulated multi-GPU run are
// engines of the same process, one host thread each)
struct cudaIpcMemHandle_t { char reserved[64]; };
constexpr unsigned cudaIpcMemLazyEnablePeerAccess = 1;
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }


// ---- NCCL inside one process: the ranks are host threads (tools/emu/multi.py). Only what ssb_engine.cu calls: grouped
// ncclSend / ncclRecv (eager: a send copies its payload into the world's queue, a receive waits for it), ncclAllReduce of
// one u64 with min, the unique id (= the address of the world object). A rank that stopped with a device error releases
// the waits of the others (g_emu_peer_failed). ----
namespace emu_nccl {
struct World {
  std::mutex m;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<std::vector<unsigned char>>> box;      // (src, dst) -> messages in order
  std::vector<unsigned long long> vals;
  int arrived = 0, generation = 0;
  unsigned long long result = 0;
};
struct Comm { World* w; int rank, nranks; };
struct Op { bool send; void* buf; size_t bytes; int peer; Comm* c; };
inline thread_local std::vector<Op> g_ops;
inline thread_local int g_depth = 0;
inline size_t dtype_bytes(int t) { return t == 1 ? 1 : (t == 3 ? 4 : 8); }
inline int run(const Op& o) {
  World* w = o.c->w;
  std::unique_lock<std::mutex> lk(w->m);
  if (o.send) {
    const unsigned char* p = (const unsigned char*)o.buf;
    w->box[{o.c->rank, o.peer}].emplace_back(p, p + o.bytes);
    w->cv.notify_all();
    return 0;
  }
  auto& q = w->box[{o.peer, o.c->rank}];
  while (q.empty()) {
    if (g_emu_peer_failed.load()) return 1;
    w->cv.wait_for(lk, std::chrono::milliseconds(20));
  }
  if (q.front().size() != o.bytes) return 2;
  std::memcpy(o.buf, q.front().data(), o.bytes);
  q.pop_front();
  return 0;
}
inline int GetUniqueId(void* id128) { World* w = new World(); std::memset(id128, 0, 128); std::memcpy(id128, &w, sizeof(w)); return 0; }
template <class Id> inline int CommInitRank(void** comm, int nranks, Id id, int rank) {
  World* w; std::memcpy(&w, &id, sizeof(w));
  { std::lock_guard<std::mutex> g(w->m); w->vals.resize((size_t)nranks); }
  *comm = new Comm{w, rank, nranks};
  return 0;
}
inline int CommDestroy(void* c) { delete (Comm*)c; return 0; }      // (the world object stays: other ranks may still hold it)
inline int GroupStart() { ++g_depth; return 0; }
inline int GroupEnd() {
  if (--g_depth > 0) return 0;
  int rc = 0;
  for (const Op& o : g_ops) if (o.send) { int r = run(o); if (r && !rc) rc = r; }      // sends first: never blocks
  for (const Op& o : g_ops) if (!o.send) { int r = run(o); if (r && !rc) rc = r; }
  g_ops.clear();
  return rc;
}
inline int Send(const void* buf, size_t count, int dtype, int peer, void* c, cudaStream_t) {
  Op o{true, const_cast<void*>(buf), count * dtype_bytes(dtype), peer, (Comm*)c};
  if (g_depth) { g_ops.push_back(o); return 0; }
  return run(o);
}
inline int Recv(void* buf, size_t count, int dtype, int peer, void* c, cudaStream_t) {
  Op o{false, buf, count * dtype_bytes(dtype), peer, (Comm*)c};
  if (g_depth) { g_ops.push_back(o); return 0; }
  return run(o);
}
// one u64, min (what the GVT reduction needs)
inline int AllReduce(const void* in, void* out, size_t count, int, int, void* cc, cudaStream_t) {
  if (count != 1) return 3;
  Comm* c = (Comm*)cc;
  World* w = c->w;
  std::unique_lock<std::mutex> lk(w->m);
  w->vals[(size_t)c->rank] = *(const unsigned long long*)in;
  const int gen = w->generation;
  if (++w->arrived == c->nranks) {
    unsigned long long m = ~0ull;
    for (unsigned long long v : w->vals) m = v < m ? v : m;
    w->result = m; w->arrived = 0; ++w->generation;
    w->cv.notify_all();
  } else {
    while (w->generation == gen) {
      if (g_emu_peer_failed.load()) return 1;
      w->cv.wait_for(lk, std::chrono::milliseconds(20));
    }
  }
  *(unsigned long long*)out = w->result;
  return 0;
}
inline const char* GetErrorString(int rc) { return rc == 1 ? "a peer rank stopped" : (rc == 2 ? "message size mismatch" : "emulated NCCL error"); }
}  // namespace emu_nccl
