// ssb_engine.cu — host side of libscalesim_b200.so: owns the device memory, enqueues the per-round kernel
// sequence on one CUDA stream (or launches it as one CUDA graph), and implements the C ABI of include/scalesim_b200.h.
// There is no CPU fallback: every entry point that does work needs a CUDA device.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <chrono>
#include <functional>
#include <thread>
#include <string>
#include <vector>

#include "../../include/scalesim_b200.h"
#include "ssb_host_internal.h"
#include "ssb_kernels.cuh"

// kernel launch: <<<>>> on the device; the developer-only host emulation (tools/emu) calls the kernel body once
#ifdef SSB_EMU
#define SSB_LAUNCH(kern, grid, block, smem, stream, ...) do { (void)(grid); (void)(block); (void)(smem); (void)(stream); kern(__VA_ARGS__); } while (0)
#else
#define SSB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

using namespace ssb;

namespace {

thread_local std::string g_create_error;

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t err__ = (call);                                                          \
    if (err__ != cudaSuccess) {                                                          \
      set_error(std::string(#call) + ": " + cudaGetErrorString(err__));                  \
      return SSB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// ---- NCCL, resolved at run time so that single-GPU use has no NCCL dependency ----
struct Nccl {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* Comm;
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  static constexpr int kUint8 = 1, kUint32 = 3, kUint64 = 5, kMin = 3;
  bool load(std::string& why) {
    if (lib) return true;
#ifdef SSB_EMU
    // host emulation: the ranks are threads of one process (cuda_emu.h, emu_nccl)
    (void)why;
    lib = this;
    GetUniqueId = [](UniqueId* id) { return emu_nccl::GetUniqueId(id); };
    CommInitRank = [](Comm* c, int n, UniqueId id, int r) { return emu_nccl::CommInitRank(c, n, id, r); };
    CommDestroy = emu_nccl::CommDestroy; AllReduce = emu_nccl::AllReduce; Send = emu_nccl::Send; Recv = emu_nccl::Recv;
    GroupStart = emu_nccl::GroupStart; GroupEnd = emu_nccl::GroupEnd; GetErrorString = emu_nccl::GetErrorString;
    return true;
#endif
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { why = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) { why = "missing " name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
  }
};
Nccl g_nccl;

template <class T> cudaError_t dalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)); }

}  // namespace

struct ssb_engine {
  ssb_config cfg;
  std::string error;
  View v;
  Ctl* h_ctl = nullptr;          // pinned host mirror
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int n_sm = 148;
  u32 n_chunks = 0;
  size_t route_smem = 0;
  int g_exec = 8, g_commit = 4, g_fix = 8;      // blocks per SM of the per-event kernels (SSB_GRID=exec,commit,fix)
  int g_post = 8, g_count = 8, g_scatter = 8, g_turn = 8;      // ... of k_post and the irregular-turn kernels (g_fix unless sized by occupancy)
  int route_bps = 4;             // resident k_route blocks per SM (shared-memory bound)
  int sync_interval = 1;         // local windows per global synchronisation (peer-memory exchange, graph loop)
  int loop_unroll = 8;           // rounds per iteration of the graph's WHILE node (SSB_UNROLL)
  std::vector<void*> allocs;
  // partition (host copies)
  std::vector<u32> h_owner, h_local;
  bool partition_set = false;
  // model data
  u32* d_phold_table = nullptr; u32 phold_table_size = 0;
  long long h_phold_params[2] = {10000, 0};
  std::vector<long long> h_lat; std::vector<int> h_rem;
  bool have_lat = false, have_rem = false;
  u32* d_route = nullptr; u32 route_len = 0;
  u32* d_roads = nullptr; u32 roads_len = 0;
  // staging
  HostEvent* d_stage = nullptr;
  long long* d_unpack = nullptr; size_t unpack_cap = 0;
  void* d_pack[2] = {nullptr, nullptr};      // staging halves of ssb_drain_committed_packed
  cudaStream_t copy_stream = nullptr;
  u64* h_poll = nullptr;                     // pinned: progress of the committed log while the loop graph runs
  char* h_stage[2] = {nullptr, nullptr};     // pinned staging halves of the streaming drains (ssb_run_stream_*)
  cudaEvent_t pack_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
  u64 log_drained = 0;
  double loop_ms = 0.0;
  long long rounds = 0;
  u64* d_inv = nullptr;                      // exact-differential repeat: per-LP invalid-from keys (View::inv while a history is loaded)
  // comm
  Nccl::Comm comm = nullptr;
  void* mbox_block = nullptr;              // own mailbox (exported with CUDA IPC)
  std::vector<void*> peer_blocks;          // peers' mailboxes, opened with CUDA IPC
  // mailbox block layout (see View): header | counts | GVT slots | records
  size_t mbox_cnt_bytes() const { return 256 + (size_t)2 * cfg.world * 2 * 128; }
  size_t mbox_gvt_bytes() const { return ((size_t)2 * cfg.world * 8 + 255) & ~(size_t)255; }
  size_t mbox_bytes() const {
    return mbox_cnt_bytes() + mbox_gvt_bytes() + (size_t)2 * cfg.world * ((size_t)v.cap_send + v.cap_anti) * sizeof(Ev);
  }
  // initial states (for ssb_reset)
  u32* d_state0 = nullptr;
  // kernel timing
  enum { K_MIN, K_PLAN, K_COMMIT, K_EXEC, K_HIST_FLAG, K_FIX_COUNT, K_FIX_SEGS, K_FIX_SCATTER, K_FIX_SORT, K_FIX_TURN,
         K_ROUTE_ANTI, K_ROUTE, K_ROUTE_RECV, K_POST, K_COUNT };
  struct Timed { int id; cudaEvent_t a, b; };
  std::vector<Timed> timed;
  std::vector<cudaEvent_t> event_pool;
  double k_ms[K_COUNT] = {0};
  long long k_launches[K_COUNT] = {0};
  bool timing() const { return (cfg.flags & SSB_FLAG_KERNEL_TIMING) != 0; }
  void t_begin(int id) {
    if (!timing()) return;
    Timed t; t.id = id;
    cudaEvent_t* ev[2] = {&t.a, &t.b};
    for (auto p : ev) {
      if (!event_pool.empty()) { *p = event_pool.back(); event_pool.pop_back(); }
      else cudaEventCreate(p);
    }
    cudaEventRecord(t.a, stream);
    timed.push_back(t);
  }
  void t_end() { if (timing()) cudaEventRecord(timed.back().b, stream); }
  void t_collect() {
    for (auto& t : timed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) { k_ms[t.id] += ms; k_launches[t.id]++; }
      event_pool.push_back(t.a); event_pool.push_back(t.b);
    }
    timed.clear();
  }
  int reset_sim();
  // whole window loop as ONE CUDA graph: WHILE(!done) { round }, with an IF node around the irregular-turn kernels
#ifndef SSB_EMU
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
#endif
  View vg;                       // view baked into the graph's kernel nodes (use_graph = 1 + the conditional handles)
  bool graph_dirty = true;
  bool graph_enabled = true;
  int build_graph();
  void destroy_graph();
#ifdef SSB_EMU
  int run_graph_emulated();
#endif

  void set_error(const std::string& s) { error = s; }

  PholdModel::Data phold_data() const {
    PholdModel::Data d;
    d.table = d_phold_table; d.table_size = phold_table_size;
    d.lookahead = (u32)h_phold_params[0]; d.num_lp = (u32)h_phold_params[1];
    return d;
  }
  TrafficModel::Data traffic_data() const {
    TrafficModel::Data d;
    d.route = d_route; d.route_len = route_len; d.roads = d_roads; d.roads_len = roads_len;
    return d;
  }

  // the ONE place that maps ssb_config.model to a device twin: f(model tag, its data)
  template <class F> void with_model(F&& f) const {
    switch (cfg.model) {
      case SSB_MODEL_PHOLD: f(PholdModel{}, phold_data()); break;
      case SSB_MODEL_PHOLD_STATEFUL: f(PholdStatefulModel{}, phold_data()); break;
      case SSB_MODEL_TRAFFIC: f(TrafficModel{}, traffic_data()); break;
    }
  }

  int grid(size_t n, int threads) const {
    size_t b = (n + threads - 1) / threads;
    size_t cap = (size_t)n_sm * 8;
    return (int)std::max<size_t>(1, std::min(b, cap));
  }

  template <class T> int alloc(T** p, size_t n) {
    CK(dalloc(p, n));
    allocs.push_back(*p);
    return SSB_OK;
  }
  int fill32(u32* p, u32 val, size_t n) {
    if (n) SSB_LAUNCH(k_fill_u32, grid(n, 256), 256, 0, stream, p, val, n);
    CK(cudaGetLastError());
    return SSB_OK;
  }
  int state_words() const {
    int w = 0;
    with_model([&](auto m, auto) { w = decltype(m)::kStateWords; });
    return w;
  }

  int init();
  int sync_ctl();
  int check_device_error();
  int model_ready();
  u32 epoch = 0;
  int enqueue_round();
  int exchange();
  void launch_commit(const View& vv);
  void launch_exec(const View& vv);
  void launch_fix(const View& vv, bool timed_);
  void launch_route(const View& vv, int which, u32 peer, u32 filter);
  u32 owner_of(long long lp) const { return h_owner.empty() ? (u32)(lp % cfg.world) : h_owner[(size_t)lp]; }
  u32 local_of(long long lp) const { return h_local.empty() ? (u32)(lp / cfg.world) : h_local[(size_t)lp]; }
};

int ssb_engine::init() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) { set_error("no CUDA device"); return SSB_ERR_CUDA; }
  CK(cudaSetDevice(cfg.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg.device));
  n_sm = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&ev0));
  CK(cudaEventCreate(&ev1));
  CK(cudaHostAlloc(reinterpret_cast<void**>(&h_ctl), sizeof(Ctl), cudaHostAllocDefault));

  std::memset(&v, 0, sizeof(v));
  v.rank = (u32)cfg.rank; v.world = (u32)cfg.world;
  v.n_lp = (u32)cfg.n_lp;
  v.n_lp_local = (u32)((cfg.n_lp - cfg.rank + cfg.world - 1) / cfg.world);   // default modulo partition
  v.part_mode = 0;
  v.bucket_ticks = (u32)cfg.bucket_ticks;
  v.finish_time = (u32)cfg.finish_time;
  v.window = (u32)cfg.window_buckets;
  v.nb = (u32)((cfg.finish_time + cfg.bucket_ticks - 1) / cfg.bucket_ticks);
  if (v.nb == 0) v.nb = 1;
  // chunk size: as large as possible while the per-bucket tail waste (one partly filled chunk per bucket) stays
  // below half of the pool
  v.ch_log = 12;
  while (v.ch_log > 4 && ((size_t)v.nb << v.ch_log) > (size_t)cfg.max_events / 2) --v.ch_log;
  const u32 chs = 1u << v.ch_log;
  n_chunks = (u32)((cfg.max_events + chs - 1) / chs) + v.nb + 1;   // + tail waste + the sacrificial chunk 0
  if ((size_t)v.nb * n_chunks * sizeof(u32) > ((size_t)4 << 30)) {
    set_error("calendar chunk table would exceed 4 GiB: raise bucket_ticks");
    return SSB_ERR_INVALID;
  }
  v.maxc = n_chunks;
  v.cap_batch = (u32)cfg.max_batch;
  v.cap_aux = std::max<u32>(1024, v.cap_batch);      // anti-messages, re-execution outputs and re-queued events of a round
  v.cap_fix = v.cap_batch + v.cap_aux;               // entries of the irregular LPs' segments (arrivals + undone events)
  v.cap_snap = (u32)std::max<int64_t>(cfg.max_snapshots, 1024);
  v.cap_send = cfg.world > 1 ? std::max<u32>(v.cap_batch / 4, 1u << 16) : 1;
  v.cap_anti = cfg.world > 1 ? std::min<u32>(v.cap_send, v.cap_aux) : 1;
  v.cap_log = (u64)cfg.max_committed;
  v.state_words = (u32)state_words();
  v.hash_hi = 4u;
  v.hash_cap = 1u << 21;
  if (const char* h = getenv("SSB_HASH_HI")) v.hash_hi = (u32)atoi(h);
  if (const char* h = getenv("SSB_HASH_CAP_LOG")) v.hash_cap = 1u << atoi(h);
  u32 hcap = 1024;
  while (hcap < 2u * v.cap_batch || (hcap < v.hash_hi * v.cap_batch && hcap < v.hash_cap)) hcap <<= 1;
  v.hmask = hcap - 1u;
  if (getenv("SSB_NO_GRAPH")) graph_enabled = false;
  if (const char* u = getenv("SSB_UNROLL")) { int x = atoi(u); if (x >= 1 && x <= 16) loop_unroll = x; }
  if (const char* gs = getenv("SSB_GRID")) {
    int a = 0, b = 0, c2 = 0;
    if (sscanf(gs, "%d,%d,%d", &a, &b, &c2) == 3 && a > 0 && b > 0 && c2 > 0) { g_exec = a; g_commit = b; g_fix = c2; }
  }

  const size_t n_slots = (size_t)n_chunks << v.ch_log;
  int rc;
  if ((rc = alloc(&v.ev, n_slots))) return rc;
  if ((rc = alloc(&v.meta, n_slots))) return rc;
  if ((rc = alloc(&v.pflag, n_slots))) return rc;
  if ((rc = alloc(&v.chunk_tab, (size_t)v.nb * v.maxc))) return rc;
  if ((rc = alloc(&v.free_stack, n_chunks))) return rc;
  if ((rc = alloc(&v.fill, v.nb))) return rc;
  if ((rc = alloc(&v.consumed, v.nb))) return rc;
  if ((rc = alloc(&v.lpw, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.state, (size_t)v.n_lp_local * v.state_words))) return rc;
  if ((rc = alloc(&v.hset, hcap))) return rc;
  if ((rc = alloc(&v.fixlp, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.fixcnt, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.fixstart, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.fixpos, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.fixund, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.fixmin, v.n_lp_local))) return rc;
  if ((rc = alloc(&v.sorted, v.cap_fix))) return rc;
  if ((rc = alloc(&v.ent_out, v.cap_fix))) return rc;
  if ((rc = alloc(&v.outbox, (size_t)v.cap_batch + v.cap_aux))) return rc;
  if ((rc = alloc(&v.anti_box, v.cap_aux))) return rc;
  if ((rc = alloc(&v.gather, std::max<u32>(v.window, 1) + 4 * 64 + 2))) return rc;      // + the lead of a synchronisation interval (k_plan: kLeadMax)
  if ((rc = alloc(&v.scan, v.nb))) return rc;
  if ((rc = alloc(&v.commit, v.nb))) return rc;
  if ((rc = alloc(&v.sendbuf, (size_t)v.world * v.cap_send))) return rc;
  if ((rc = alloc(&v.recvbuf, (size_t)v.world * v.cap_send))) return rc;
  if ((rc = alloc(&v.log, (size_t)v.cap_log))) return rc;
  if ((cfg.flags & SSB_FLAG_HISTORY) && v.cap_log && (rc = alloc(&v.log_sent, (size_t)v.cap_log))) return rc;
  if ((rc = alloc(&v.snap, (size_t)v.cap_snap * v.state_words))) return rc;
  if ((rc = alloc(&v.ctl, 1))) return rc;
  if ((rc = alloc(&d_stage, v.cap_batch))) return rc;

  if ((rc = alloc(&d_state0, (size_t)v.n_lp_local * v.state_words))) return rc;
  CK(cudaMemsetAsync(v.state, 0, sizeof(u32) * (size_t)v.n_lp_local * v.state_words, stream));
  CK(cudaMemsetAsync(d_state0, 0, sizeof(u32) * (size_t)v.n_lp_local * v.state_words, stream));
  if ((rc = reset_sim())) return rc;
  v.route_slots = std::min<u32>(v.nb, 2048);
  route_smem = (size_t)kRouteTile * sizeof(Ev) + (size_t)(v.route_slots + 1 + v.world) * 16;
#ifndef SSB_EMU
  CK(cudaFuncSetAttribute(k_route, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)route_smem));
  CK(cudaFuncSetAttribute(k_route, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&route_bps, k_route, kRouteThreads, route_smem));
#endif
  if (route_bps < 1) route_bps = 1;
#ifndef SSB_EMU
  if (!getenv("SSB_GRID")) {
    // persistent grids of exactly one wave: resident blocks per SM of the model's instantiation (a grid sized for
    // more blocks than fit leaves a nearly empty second wave: measured +20 % on k_exec at 33 registers)
    int be = g_exec, bc = g_commit, bt = g_fix, bp = g_fix, bn = g_fix, bs = g_fix;
    with_model([&](auto m, auto) {
      typedef decltype(m) M;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&be, k_exec<M>, 256, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bc, k_commit<M>, (int)kCommitThreads, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt, k_fix_turn<M>, 256, 0);
    });
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bp, k_post, 256, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bn, k_fix_count, 256, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bs, k_fix_scatter, 256, 0);
    if (be >= 1) g_exec = be;
    if (bc >= 1) g_commit = bc;
    if (bt >= 1) g_turn = bt;
    if (bp >= 1) g_post = bp;
    if (bn >= 1) g_count = bn;
    if (bs >= 1) g_scatter = bs;
  } else {
    g_post = g_count = g_scatter = g_turn = g_fix;
  }
#endif
  if (getenv("SSB_VERBOSE")) fprintf(stderr, "[ssb] blocks per SM: k_exec %d, k_commit %d, k_post %d, k_fix_count %d, k_fix_scatter %d, k_fix_turn %d\n",
                                     g_exec, g_commit, g_post, g_count, g_scatter, g_turn);
  if (getenv("SSB_VERBOSE")) fprintf(stderr, "[ssb] k_route: %zu B shared memory per block, %d blocks per SM\n", route_smem, route_bps);
  if (const char* rb = getenv("SSB_ROUTE_BPS")) { int x = atoi(rb); if (x >= 1 && x <= route_bps) route_bps = x; }
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  return SSB_OK;
}

int ssb_engine::reset_sim() {
  Ctl init;
  std::memset(&init, 0, sizeof(init));
  init.local_min = KEY_MAX; init.gvt = 0; init.prev_gvt = KEY_MAX;
  init.epoch = ++epoch;
  init.free_top = n_chunks - 1; init.free_low = n_chunks - 1;
  CK(cudaStreamSynchronize(stream));
  CK(cudaMemcpyAsync(v.ctl, &init, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
  CK(cudaStreamSynchronize(stream));   // `init` lives on the stack
  SSB_LAUNCH(k_iota_u32, grid(n_chunks - 1, 256), 256, 0, stream, v.free_stack, 1u, (size_t)(n_chunks - 1));   // chunks 1..n-1 are free
  int rc;
  if ((rc = fill32(v.chunk_tab, NIL, (size_t)v.nb * v.maxc))) return rc;
  CK(cudaMemsetAsync(v.fill, 0, sizeof(u32) * v.nb, stream));
  CK(cudaMemsetAsync(v.consumed, 0, sizeof(u32) * v.nb, stream));
  CK(cudaMemsetAsync(v.lpw, 0, sizeof(uint2) * (size_t)v.n_lp_local, stream));
  CK(cudaMemsetAsync(v.hset, 0, sizeof(u32) * ((size_t)v.hmask + 1), stream));
  CK(cudaMemcpyAsync(v.state, d_state0, sizeof(u32) * (size_t)v.n_lp_local * v.state_words, cudaMemcpyDeviceToDevice, stream));
  if (v.inv) { v.inv = nullptr; graph_dirty = true; }      // a loaded history goes with the reset
  if (mbox_block) {
    // own mailbox: counts and GVT slots of the finished run (all partitions must reset together, before any of them
    // starts the next run)
    CK(cudaMemsetAsync(static_cast<char*>(mbox_block) + 256, 0, mbox_cnt_bytes() - 256 + mbox_gvt_bytes(), stream));
  }
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  log_drained = 0;
  loop_ms = 0.0;
  rounds = 0;
  t_collect();
  for (int i = 0; i < K_COUNT; ++i) { k_ms[i] = 0; k_launches[i] = 0; }
  return SSB_OK;
}

int ssb_engine::sync_ctl() {
  CK(cudaMemcpyAsync(h_ctl, v.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  return SSB_OK;
}

int ssb_engine::check_device_error() {
  u32 e = h_ctl->error;
  if (!e) return SSB_OK;
#ifdef SSB_EMU
  g_emu_peer_failed.store(1);      // (tools/emu/multi.py: the peer partitions stop waiting for this one)
#endif
  std::string s = "device error:";
  if (e & E_POOL) s += " event pool exhausted (raise max_events)";
  if (e & E_RANGE) s += " value out of device record range / destination not owned";
  if (e & E_BATCH) s += " rollback outbox / segment overflow (raise max_batch)";
  if (e & E_SNAP) s += " snapshot ring overflow (raise max_snapshots)";
  if (e & E_HORIZON) s += " calendar horizon exceeded";
  if (e & E_SENDBUF) s += " send buffer overflow (raise max_batch)";
  // (after a failed GVT exchange the rounds that were already enqueued run on an unsynchronised calendar and trip the
  // invariant checks: a consequence, not a cause — only reported when the exchange itself was fine)
  if ((e & E_INTERNAL) && !(e & E_COMM)) s += " internal invariant violated (#" + std::to_string(h_ctl->where) + ", " + std::to_string(h_ctl->where_arg) + ")";
  if (e & E_STALL) s += " no GVT progress in 100000 rounds";
  if (e & E_COMM) s += " a peer partition did not reach the GVT exchange";
  set_error(s);
  if (e & (E_POOL | E_BATCH | E_SNAP | E_SENDBUF)) return SSB_ERR_CAPACITY;
  if (e & E_RANGE) return SSB_ERR_RANGE;
  if (e & E_COMM) return SSB_ERR_COMM;
  return SSB_ERR_INVALID;
}

int ssb_engine::model_ready() {
  if (cfg.model == SSB_MODEL_PHOLD || cfg.model == SSB_MODEL_PHOLD_STATEFUL) {
    if (!d_phold_table) {
      if (!(have_lat && have_rem)) { set_error("PHOLD model data slots 0 and 1 not set"); return SSB_ERR_MODEL; }
      u32 n = (u32)h_lat.size();
      if (h_rem.size() != h_lat.size()) { set_error("PHOLD tables differ in length"); return SSB_ERR_MODEL; }
      long long* d_lat = nullptr; int* d_rem = nullptr;
      CK(dalloc(&d_lat, n)); CK(dalloc(&d_rem, n));
      int rc = alloc(&d_phold_table, n);
      if (rc) return rc;
      CK(cudaMemcpyAsync(d_lat, h_lat.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(d_rem, h_rem.data(), sizeof(int) * n, cudaMemcpyHostToDevice, stream));
      SSB_LAUNCH(k_pack_phold_table, grid(n, 256), 256, 0, stream, (const long long*)d_lat, (const int*)d_rem, d_phold_table, n, v.ctl);
      CK(cudaStreamSynchronize(stream));
      CK(cudaFree(d_lat)); CK(cudaFree(d_rem));
      phold_table_size = n;
      h_lat.clear(); h_lat.shrink_to_fit(); h_rem.clear(); h_rem.shrink_to_fit();
    }
    if (h_phold_params[1] <= 0) h_phold_params[1] = cfg.n_lp;
  } else if (cfg.model == SSB_MODEL_TRAFFIC) {
    if (!d_route) { set_error("TrafficSim model data slot 0 (route pool) not set"); return SSB_ERR_MODEL; }
  }
  return SSB_OK;
}

void ssb_engine::launch_route(const View& vv, int which, u32 peer, u32 filter) {
  t_begin(which == 0 ? K_ROUTE_ANTI : (which == 2 ? K_ROUTE_RECV : K_ROUTE));
  SSB_LAUNCH(k_route, n_sm * route_bps, kRouteThreads, route_smem, stream, vv, which, peer, filter);
  t_end();
}

void ssb_engine::launch_commit(const View& vv) {
  const int g = n_sm * g_commit;
  with_model([&](auto m, auto md) { SSB_LAUNCH(k_commit<decltype(m)>, g, kCommitThreads, 0, stream, vv, md); });
}

void ssb_engine::launch_exec(const View& vv) {
  const int g = n_sm * g_exec;
  with_model([&](auto m, auto md) { SSB_LAUNCH(k_exec<decltype(m)>, g, 256, 0, stream, vv, md); });
}

// the irregular turns of a round (inside the graph: the body of an IF node)
void ssb_engine::launch_fix(const View& vv, bool timed_) {
  auto tb = [&](int id) { if (timed_) t_begin(id); };
  auto te = [&]() { if (timed_) t_end(); };
  if (vv.inv) { tb(K_HIST_FLAG); SSB_LAUNCH(k_hist_flag, n_sm * g_count, 256, 0, stream, vv); te(); }
  tb(K_FIX_COUNT); SSB_LAUNCH(k_fix_count, n_sm * g_count, 256, 0, stream, vv); te();
  tb(K_FIX_SEGS); SSB_LAUNCH(k_fix_segs, n_sm, 256, 0, stream, vv); te();
  tb(K_FIX_SCATTER); SSB_LAUNCH(k_fix_scatter, n_sm * g_scatter, 256, 0, stream, vv); te();
  tb(K_FIX_SORT); SSB_LAUNCH(k_fix_sort, n_sm * 4, kSortThreads, 0, stream, vv); te();
  tb(K_FIX_TURN);
  with_model([&](auto m, auto md) { SSB_LAUNCH(k_fix_turn<decltype(m)>, n_sm * g_turn, 256, 0, stream, vv, md); });
  te();
  launch_route(vv, 0, 0, 0);
}

// counts all-to-all, then payload all-to-all (grouped ncclSend/ncclRecv), then integrate what arrived
int ssb_engine::exchange() {
  if (cfg.world <= 1 || v.p2p) return SSB_OK;      // peer-memory exchange: the receive side runs at the next round's start
  if (!comm) { set_error("world > 1 but ssb_comm_init was not called"); return SSB_ERR_COMM; }
  Nccl& n = g_nccl;
  int rc = 0, first_rc = 0;
  std::string first_call;
  // inside a group every call is still issued after a failure and the group is always closed: the peers' matching
  // sends / receives must not be left hanging
#define NC(call) do { rc = (call); if (rc != 0 && !first_rc) { first_rc = rc; first_call = #call; } } while (0)
  NC(n.GroupStart());
  for (int r = 0; r < cfg.world; ++r) {
    if (r == cfg.rank) continue;
    NC(n.Send(&v.ctl->send_count[r], 1, Nccl::kUint32, r, comm, stream));
    NC(n.Recv(&v.ctl->recv_count[r], 1, Nccl::kUint32, r, comm, stream));
  }
  NC(n.GroupEnd());
  if (first_rc) { set_error(first_call + ": " + n.GetErrorString(first_rc)); return SSB_ERR_COMM; }
  int s = sync_ctl();
  if (s) return s;
  // both ends of a transfer see the same unclamped count (our send count is the peer's receive count), so a transfer
  // that does not fit is skipped by both, and both report the overflow
  bool overflow = false;
  NC(n.GroupStart());
  for (int r = 0; r < cfg.world; ++r) {
    if (r == cfg.rank) continue;
    const u32 ns = h_ctl->send_count[r], nr = h_ctl->recv_count[r];
    if (ns > v.cap_send || nr > v.cap_send) overflow = true;
    if (ns && ns <= v.cap_send) NC(n.Send(v.sendbuf + (size_t)r * v.cap_send, (size_t)ns * sizeof(Ev), Nccl::kUint8, r, comm, stream));
    if (nr && nr <= v.cap_send) NC(n.Recv(v.recvbuf + (size_t)r * v.cap_send, (size_t)nr * sizeof(Ev), Nccl::kUint8, r, comm, stream));
  }
  NC(n.GroupEnd());
#undef NC
  if (first_rc) { set_error(first_call + ": " + n.GetErrorString(first_rc)); return SSB_ERR_COMM; }
  if (overflow) { set_error("send / receive buffer overflow (raise max_batch)"); return SSB_ERR_CAPACITY; }
  for (int r = 0; r < cfg.world; ++r)
    if (r != cfg.rank && h_ctl->recv_count[r]) launch_route(v, 2, (u32)r, 0);
  CK(cudaGetLastError());
  return SSB_OK;
}

int ssb_engine::enqueue_round() {
  if (cfg.world > 1) {
    if (v.p2p) {
      // GVT candidates cross through the mailboxes; when all have arrived, what the peers appended to this GPU's
      // mailbox in the previous round is complete: integrate it (anti-messages first) before planning the window
      t_begin(K_MIN);
      SSB_LAUNCH(k_gvt, 1, 256, 0, stream, v);
      t_end();
      launch_route(v, 2, 0, 0);
    } else {
      if (!comm) { set_error("world > 1 but ssb_comm_init was not called"); return SSB_ERR_COMM; }
      t_begin(K_MIN);
      SSB_LAUNCH(k_min, 1, 1024, 0, stream, v);
      t_end();
      int rc = g_nccl.AllReduce(&v.ctl->local_min, &v.ctl->gvt, 1, Nccl::kUint64, Nccl::kMin, comm, stream);
      if (rc != 0) { set_error(std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc)); return SSB_ERR_COMM; }
    }
  }
  t_begin(K_PLAN);
  SSB_LAUNCH(k_plan, 1, 1024, 0, stream, v, cfg.world > 1 ? 1 : 0);
  t_end();
  t_begin(K_COMMIT);
  launch_commit(v);
  t_end();
  t_begin(K_EXEC);
  launch_exec(v);
  t_end();
  launch_fix(v, true);
  launch_route(v, 1, 0, 0);
  t_begin(K_POST);
  SSB_LAUNCH(k_post, n_sm * g_post, 256, 0, stream, v);
  t_end();
  CK(cudaGetLastError());
  ++rounds;
  return exchange();
}

// ------------------------------------------------------------------------------------------------
// The window loop as one CUDA graph: a WHILE node whose condition k_plan keeps at "GVT has not reached finish_time
// and no device error", and inside it one round. The commit pass runs as a parallel branch next to k_exec (they
// touch disjoint buckets; both join before k_route, which is the only other user of the chunk free stack). The
// kernels that only irregular turns need sit in an IF node whose condition the last block of k_exec sets on the
// device. The host launches the graph once per ssb_run and is not involved in the loop at all — also for N > 1,
// because GVT and the barrier go through peer memory.
// ------------------------------------------------------------------------------------------------
#ifdef SSB_EMU
// Host emulation: the graph below is interpreted — same nodes in the same order, the WHILE / IF conditions read from
// the table cudaGraphSetConditional writes to (cuda_emu.h). The kernels run with use_graph = 1, so what they do only on
// the graph path (loop condition and stall detection in k_plan, the IF condition set by the last block of k_exec,
// several windows per synchronisation) is exercised where there is no GPU.
void ssb_engine::destroy_graph() {}
int ssb_engine::build_graph() { graph_dirty = false; return SSB_OK; }
int ssb_engine::run_graph_emulated() {
  vg = v;
  vg.use_graph = 1;
  vg.sync = cfg.world > 1 ? (u32)sync_interval : 1u;
  vg.h_loop = 1; vg.h_fix = 2;
  const int subs = cfg.world > 1 ? sync_interval : 1;
  g_emu_cond[1] = 1;                                  // cudaGraphCondAssignDefault, default value 1
  while (g_emu_cond[1]) {                             // WHILE node
    for (int rep = 0; rep < loop_unroll; ++rep) {     // its body: `unroll` rounds back to back
      for (int j = 0; j < subs; ++j) {
        View vj = vg;
        vj.sub = (u32)j;
        g_emu_cond[2] = 0;                            // this round's IF handle, default value 0
        if (cfg.world > 1 && j == 0) {
          SSB_LAUNCH(k_gvt, 1, 256, 0, stream, vj);
          launch_route(vj, 2, 0, 0);
        }
        SSB_LAUNCH(k_plan, 1, 1024, 0, stream, vj, cfg.world > 1 ? 1 : 0);
        if (j == 0) launch_commit(vj);
        launch_exec(vj);
        if (g_emu_cond[2]) launch_fix(vj, false);     // IF node
        launch_route(vj, 1, 0, 0);
        SSB_LAUNCH(k_post, n_sm * g_post, 256, 0, stream, vj);
      }
    }
  }
  return SSB_OK;
}
#else
void ssb_engine::destroy_graph() {
  if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
  if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
  graph_dirty = true;
}

int ssb_engine::build_graph() {
  destroy_graph();
  CK(cudaStreamSynchronize(stream));
  CK(cudaGraphCreate(&graph, 0));
  vg = v;
  vg.use_graph = 1;
  vg.sync = cfg.world > 1 ? (u32)sync_interval : 1u;
  CK(cudaGraphConditionalHandleCreate(&vg.h_loop, graph, 1, cudaGraphCondAssignDefault));
  auto add_cond = [&](cudaGraph_t parent, cudaGraphConditionalHandle h, cudaGraphConditionalNodeType type,
                      const std::vector<cudaGraphNode_t>& deps, cudaGraphNode_t* node, cudaGraph_t* body) -> cudaError_t {
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = type;
    p.conditional.size = 1;
    cudaError_t err = cudaGraphAddNode(node, parent, deps.empty() ? nullptr : deps.data(), deps.size(), &p);
    if (err == cudaSuccess) *body = p.conditional.phGraph_out[0];
    return err;
  };
  // captures what `fn` enqueues on `stream` into graph `g` behind `deps`; returns the tail nodes
  auto capture = [&](cudaGraph_t g, const std::vector<cudaGraphNode_t>& deps, std::vector<cudaGraphNode_t>* tail,
                     const std::function<void()>& fn) -> cudaError_t {
    cudaError_t err = cudaStreamBeginCaptureToGraph(stream, g, deps.empty() ? nullptr : deps.data(), nullptr, deps.size(),
                                                    cudaStreamCaptureModeThreadLocal);
    if (err != cudaSuccess) return err;
    fn();
    cudaStreamCaptureStatus st;
    const cudaGraphNode_t* d = nullptr;
    size_t nd = 0;
    err = cudaStreamGetCaptureInfo(stream, &st, nullptr, nullptr, &d, &nd);
    if (err == cudaSuccess && tail) tail->assign(d, d + nd);
    cudaGraph_t out = nullptr;
    cudaError_t err2 = cudaStreamEndCapture(stream, &out);
    return err != cudaSuccess ? err : err2;
  };
  cudaGraphNode_t n_while;
  cudaGraph_t body;
  CK(add_cond(graph, vg.h_loop, cudaGraphCondTypeWhile, {}, &n_while, &body));
  // The WHILE body holds `unroll` rounds back to back (a loop iteration costs ~4.7 us on B200, a kernel node ~0.9 us:
  // tools/mb_graph.cu): every round's k_plan writes the loop condition, and once GVT has reached finish_time the
  // remaining rounds of the iteration find nothing to do. Inside a round: [N > 1: k_gvt, k_route(recv)] k_plan,
  // then k_commit next to k_exec, the IF node of the irregular turns, k_route next to k_post; with
  // ssb_set_sync_interval further windows without a global synchronisation follow (k_plan .. k_post only).
  const int subs = cfg.world > 1 ? sync_interval : 1;
  std::vector<cudaGraphNode_t> tail;      // what the next round has to wait for
  for (int rep = 0; rep < loop_unroll; ++rep) {
    for (int j = 0; j < subs; ++j) {
      View vj = vg;
      vj.sub = (u32)j;
      CK(cudaGraphConditionalHandleCreate(&vj.h_fix, graph, 0, cudaGraphCondAssignDefault));
      std::vector<cudaGraphNode_t> t_plan, t_commit, t_exec, t_route, t_post;
      CK(capture(body, tail, &t_plan, [&] {
        if (cfg.world > 1 && j == 0) {
          // peer-memory exchange: GVT + barrier, then the receive side of the previous exchange round (one launch
          // over the anti-messages and the positives of all peers)
          SSB_LAUNCH(k_gvt, 1, 256, 0, stream, vj);
          launch_route(vj, 2, 0, 0);
        }
        SSB_LAUNCH(k_plan, 1, 1024, 0, stream, vj, cfg.world > 1 ? 1 : 0);
      }));
      std::vector<cudaGraphNode_t> join;
      if (j == 0) {      // GVT only moves at a global synchronisation: later windows have nothing new to commit
        CK(capture(body, t_plan, &t_commit, [&] { launch_commit(vj); }));
        join = t_commit;
      }
      CK(capture(body, t_plan, &t_exec, [&] { launch_exec(vj); }));
      // k_route(anti) pops chunks off the free stack that the commit pass pushes to: the IF node waits for both branches
      join.insert(join.end(), t_exec.begin(), t_exec.end());
      cudaGraphNode_t n_if;
      cudaGraph_t b_if;
      CK(add_cond(body, vj.h_fix, cudaGraphCondTypeIf, join, &n_if, &b_if));
      CK(capture(b_if, {}, nullptr, [&] { launch_fix(vj, false); }));
      CK(capture(body, {n_if}, &t_route, [&] { launch_route(vj, 1, 0, 0); }));
      CK(capture(body, {n_if}, &t_post, [&] { SSB_LAUNCH(k_post, n_sm * g_post, 256, 0, stream, vj); }));      // next to k_route
      tail = t_route;
      tail.insert(tail.end(), t_post.begin(), t_post.end());
    }
  }
  CK(cudaGetLastError());
  CK(cudaGraphInstantiate(&graph_exec, graph, 0));
  graph_dirty = false;
  return SSB_OK;
}
#endif

namespace {
template <class T> bool regrow(ssb_engine* e, T** p, size_t n) {
  T* q = nullptr;
  if (dalloc(&q, n) != cudaSuccess) return false;
  e->allocs.push_back(q);
  *p = q;
  return true;
}
}  // namespace

// ======================================= C ABI =======================================
void ssb_internal_set_error(const std::string& what) { g_create_error = what; }
void ssb_internal_set_engine_error(ssb_engine* e, const std::string& what) { if (e) e->set_error(what); else g_create_error = what; }

extern "C" {

int ssb_abi_version(void) { return SSB_ABI_VERSION; }

int ssb_get_config(ssb_engine* e, ssb_config* out) {
  if (!e || !out) return SSB_ERR_INVALID;
  *out = e->cfg;
  return SSB_OK;
}

const char* ssb_last_error(ssb_engine* e) { return e ? e->error.c_str() : g_create_error.c_str(); }

void ssb_destroy(ssb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  e->destroy_graph();
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
#ifndef SSB_EMU
  for (void* p : e->peer_blocks) if (p) cudaIpcCloseMemHandle(p);
#endif
  if (e->mbox_block) cudaFree(e->mbox_block);
  for (void* p : e->allocs) cudaFree(p);
  if (e->d_unpack) cudaFree(e->d_unpack);
  for (int i = 0; i < 2; ++i) {
    if (e->d_pack[i]) cudaFree(e->d_pack[i]);
    if (e->pack_done[i]) cudaEventDestroy(e->pack_done[i]);
    if (e->copy_done[i]) cudaEventDestroy(e->copy_done[i]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->h_poll) cudaFreeHost(e->h_poll);
  for (int i = 0; i < 2; ++i) if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]);
  if (e->h_ctl) cudaFreeHost(e->h_ctl);
  e->t_collect();
  for (cudaEvent_t x : e->event_pool) cudaEventDestroy(x);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int ssb_create(const ssb_config* cfg, ssb_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return SSB_ERR_INVALID; }
#ifdef SSB_EMU
  g_emu_peer_failed.store(0);
#endif
  *out = nullptr;
  auto bad = [&](const char* m) { g_create_error = m; return SSB_ERR_INVALID; };
  if (cfg->abi_version != SSB_ABI_VERSION) return bad("abi_version mismatch");
  if (cfg->model != SSB_MODEL_PHOLD && cfg->model != SSB_MODEL_TRAFFIC && cfg->model != SSB_MODEL_PHOLD_STATEFUL) return bad("unknown model");
  if (cfg->world < 1 || cfg->world > MAX_WORLD || cfg->rank < 0 || cfg->rank >= cfg->world) return bad("bad rank/world");
  if (cfg->n_lp < 1 || cfg->n_lp >= 0xFFFFFFFFll) return bad("n_lp out of range");
  if (cfg->finish_time < 1 || cfg->finish_time >= 0x7FFFFFFFll) return bad("finish_time out of range (must be < 2^31)");
  if (cfg->bucket_ticks < 1 || cfg->window_buckets < 1) return bad("bucket_ticks and window_buckets must be >= 1");
  if (cfg->max_events < 1 || cfg->max_events > 0xF0000000ll) return bad("max_events out of range");
  if (cfg->max_batch < 1 || cfg->max_batch >= (1ll << 28)) return bad("max_batch out of range (must be < 2^28)");
  if (cfg->max_committed < 0) return bad("max_committed out of range");
  ssb_engine* e = new ssb_engine();
  e->cfg = *cfg;
  if (e->cfg.rounds_per_sync < 1) e->cfg.rounds_per_sync = 4;
  int rc = e->init();
  if (rc != SSB_OK) {
    g_create_error = e->error;
    ssb_destroy(e);
    return rc;
  }
  *out = e;
  return SSB_OK;
}

int ssb_set_partition(ssb_engine* e, const int64_t* lp_to_part, int64_t n_lp) {
  if (!e) return SSB_ERR_INVALID;
  if (e->partition_set) { e->set_error("partition already set"); return SSB_ERR_INVALID; }
  e->graph_dirty = true;
  cudaSetDevice(e->cfg.device);
  if (!lp_to_part) { e->partition_set = true; return SSB_OK; }   // default: lp % world
  if (n_lp != e->cfg.n_lp) { e->set_error("partition length != n_lp"); return SSB_ERR_INVALID; }
  bool modulo = true;
  for (int64_t i = 0; i < n_lp; ++i) {
    if (lp_to_part[i] < 0 || lp_to_part[i] >= e->cfg.world) { e->set_error("partition value out of [0, world) (reference quirk Q12)"); return SSB_ERR_INVALID; }
    if (lp_to_part[i] != i % e->cfg.world) modulo = false;
  }
  e->partition_set = true;
  if (modulo) return SSB_OK;
  e->h_owner.resize((size_t)n_lp); e->h_local.resize((size_t)n_lp);
  std::vector<u32> counts((size_t)e->cfg.world, 0);
  for (int64_t i = 0; i < n_lp; ++i) {
    u32 o = (u32)lp_to_part[i];
    e->h_owner[(size_t)i] = o;
    e->h_local[(size_t)i] = counts[o]++;
  }
  u32 mine = counts[(size_t)e->cfg.rank];
  if (mine > e->v.n_lp_local) {
    // per-LP arrays were sized for the modulo partition: grow them
    auto& v = e->v;
    const size_t sw = v.state_words;
    if (!(regrow(e, &v.lpw, mine) && regrow(e, &v.state, mine * sw) && regrow(e, &e->d_state0, mine * sw) &&
          regrow(e, &v.fixlp, mine) && regrow(e, &v.fixcnt, mine) && regrow(e, &v.fixstart, mine) &&
          regrow(e, &v.fixpos, mine) && regrow(e, &v.fixund, mine) && regrow(e, &v.fixmin, mine))) {
      e->set_error("cudaMalloc failed"); return SSB_ERR_CUDA;
    }
    cudaMemsetAsync(e->d_state0, 0, sizeof(u32) * mine * sw, e->stream);
    cudaMemsetAsync(v.lpw, 0, sizeof(uint2) * (size_t)mine, e->stream);
    cudaMemsetAsync(v.state, 0, sizeof(u32) * mine * sw, e->stream);
  }
  e->v.n_lp_local = mine;
  u32 *d_o = nullptr, *d_l = nullptr;
  int rc;
  if ((rc = e->alloc(&d_o, (size_t)n_lp))) return rc;
  if ((rc = e->alloc(&d_l, (size_t)n_lp))) return rc;
  cudaMemcpyAsync(d_o, e->h_owner.data(), sizeof(u32) * (size_t)n_lp, cudaMemcpyHostToDevice, e->stream);
  cudaMemcpyAsync(d_l, e->h_local.data(), sizeof(u32) * (size_t)n_lp, cudaMemcpyHostToDevice, e->stream);
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SSB_ERR_CUDA; }
  e->v.part_mode = 1; e->v.part_owner = d_o; e->v.part_local = d_l;
  return SSB_OK;
}

int ssb_set_model_data(ssb_engine* e, int slot, const void* data, size_t bytes) {
  if (!e || !data) return SSB_ERR_INVALID;
  e->graph_dirty = true;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  if (e->cfg.model == SSB_MODEL_PHOLD || e->cfg.model == SSB_MODEL_PHOLD_STATEFUL) {
    if (slot == 0) { e->h_lat.assign((const long long*)data, (const long long*)data + bytes / 8); e->have_lat = true; return SSB_OK; }
    if (slot == 1) { e->h_rem.assign((const int*)data, (const int*)data + bytes / 4); e->have_rem = true; return SSB_OK; }
    if (slot == 2 && bytes == 16) { std::memcpy(e->h_phold_params, data, 16); return SSB_OK; }
    set_error("PHOLD: unknown model data slot");
    return SSB_ERR_MODEL;
  }
  if (e->cfg.model == SSB_MODEL_TRAFFIC) {
    if (slot != 0 && slot != 1) { set_error("TrafficSim: unknown model data slot"); return SSB_ERR_MODEL; }
    size_t n = bytes / 8;
    if (n >= 0xFFFFFFFFull) { set_error(slot == 0 ? "route pool too large" : "road table too large"); return SSB_ERR_RANGE; }
    if (slot == 1 && n % 3) { set_error("road table: {dst, speed, length} triples expected"); return SSB_ERR_MODEL; }
    long long* d_in = nullptr;
    CK(dalloc(&d_in, n));
    u32** dst = slot == 0 ? &e->d_route : &e->d_roads;
    int rc = e->alloc(dst, n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_in, data, n * 8, cudaMemcpyHostToDevice, e->stream));
    SSB_LAUNCH(k_narrow_u32, e->grid(n, 256), 256, 0, e->stream, (const long long*)d_in, *dst, n, e->v.ctl);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaFree(d_in));
    (slot == 0 ? e->route_len : e->roads_len) = (u32)n;
    return SSB_OK;
  }
  return SSB_ERR_MODEL;
}

int ssb_state_words(ssb_engine* e) { return e ? e->state_words() : 0; }

static int states_io(ssb_engine* e, const int64_t* lp_ids, uint32_t* words, int64_t n, bool write) {
  if (!e || !lp_ids || !words || n < 0) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  const int sw = e->state_words();
  // contiguous runs of local indices are copied with one memcpy each
  int64_t i = 0;
  while (i < n) {
    if (lp_ids[i] < 0 || lp_ids[i] >= e->cfg.n_lp || e->owner_of(lp_ids[i]) != (u32)e->cfg.rank) {
      set_error("state for an LP this rank does not own"); return SSB_ERR_INVALID;
    }
    u32 l0 = e->local_of(lp_ids[i]);
    int64_t j = i + 1;
    while (j < n && lp_ids[j] >= 0 && lp_ids[j] < e->cfg.n_lp && e->owner_of(lp_ids[j]) == (u32)e->cfg.rank &&
           e->local_of(lp_ids[j]) == l0 + (u32)(j - i)) ++j;
    size_t bytes = (size_t)(j - i) * sw * 4;
    if (write) {
      CK(cudaMemcpyAsync(e->v.state + (size_t)l0 * sw, words + (size_t)i * sw, bytes, cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(e->d_state0 + (size_t)l0 * sw, words + (size_t)i * sw, bytes, cudaMemcpyHostToDevice, e->stream));
    }
    else CK(cudaMemcpyAsync(words + (size_t)i * sw, e->v.state + (size_t)l0 * sw, bytes, cudaMemcpyDeviceToHost, e->stream));
    i = j;
  }
  CK(cudaStreamSynchronize(e->stream));
  return SSB_OK;
}
int ssb_load_states(ssb_engine* e, const int64_t* lp_ids, const uint32_t* words, int64_t n) {
  return states_io(e, lp_ids, const_cast<uint32_t*>(words), n, true);
}
int ssb_read_states(ssb_engine* e, const int64_t* lp_ids, uint32_t* words, int64_t n) {
  return states_io(e, lp_ids, words, n, false);
}

int ssb_load_events(ssb_engine* e, const ssb_event* ev, int64_t n) {
  if (!e || (!ev && n) || n < 0) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  static_assert(sizeof(ssb_event) == sizeof(HostEvent), "ssb_event layout");
  int64_t done = 0;
  while (done < n) {
    u32 m = (u32)std::min<int64_t>(n - done, e->v.cap_batch);
    CK(cudaMemcpyAsync(e->d_stage, ev + done, sizeof(HostEvent) * m, cudaMemcpyHostToDevice, e->stream));
    SSB_LAUNCH(k_pack, e->grid(m, 256), 256, 0, e->stream, e->v, (const HostEvent*)e->d_stage, m);
    e->launch_route(e->v, 3, 0, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));   // the host buffer may be reused by the caller; staging is single-buffered
    done += m;
  }
  int rc = e->sync_ctl();
  if (rc) return rc;
  return e->check_device_error();
}

int ssb_load_events_raw(ssb_engine* e, const ssb_raw_record* ev, int64_t n) {
  if (!e || (!ev && n) || n < 0) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  static_assert(sizeof(ssb_raw_record) == sizeof(Ev), "raw record layout");
  int64_t done = 0;
  while (done < n) {
    u32 m = (u32)std::min<int64_t>(n - done, e->v.cap_batch);
    // straight into the routing staging area: no unpack pass
    CK(cudaMemcpyAsync(e->v.outbox, ev + done, sizeof(Ev) * m, cudaMemcpyHostToDevice, e->stream));
    SSB_LAUNCH(k_check_raw, e->grid(m, 256), 256, 0, e->stream, e->v, m);
    e->launch_route(e->v, 3, 0, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    done += m;
  }
  int rc = e->sync_ctl();
  if (rc) return rc;
  return e->check_device_error();
}

int ssb_step(ssb_engine* e, int* done) {
  if (!e) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  int rc = e->model_ready();
  if (rc) return rc;
  if ((rc = e->enqueue_round())) return rc;
  if ((rc = e->sync_ctl())) return rc;
  if (done) *done = (int)e->h_ctl->done;
  return e->check_device_error();
}

static bool use_graph_path(const ssb_engine* e) {
  return e->graph_enabled && (e->cfg.world == 1 || e->v.p2p) && !e->timing();
}

int ssb_run(ssb_engine* e, ssb_stats* stats) {
  if (!e) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  int rc = e->model_ready();
  if (rc) return rc;
  const bool use_graph = use_graph_path(e);
  if (use_graph && e->graph_dirty && (rc = e->build_graph())) return rc;
  CK(cudaEventRecord(e->ev0, e->stream));
  u64 last_gvt = ~0ull;
  long long stalled = 0;
  if (use_graph) {
#ifndef SSB_EMU
    CK(cudaGraphLaunch(e->graph_exec, e->stream));
#else
    if ((rc = e->run_graph_emulated())) return rc;
#endif
    if ((rc = e->sync_ctl())) return rc;
    e->rounds = e->h_ctl->round;
    if ((rc = e->check_device_error())) return rc;
    if (!e->h_ctl->done) { e->set_error("window loop ended before GVT reached finish_time"); return SSB_ERR_INVALID; }
  } else for (;;) {
    for (int i = 0; i < e->cfg.rounds_per_sync; ++i)
      if ((rc = e->enqueue_round())) return rc;
    if ((rc = e->sync_ctl())) return rc;
    if ((rc = e->check_device_error())) return rc;
    if (e->h_ctl->done) break;
    // GVT must keep advancing (every round commits or integrates something); bail out instead of spinning forever
    if (e->h_ctl->gvt == last_gvt) stalled += e->cfg.rounds_per_sync; else { stalled = 0; last_gvt = e->h_ctl->gvt; }
    if (stalled > 100000) { e->set_error("no GVT progress in 100000 rounds"); return SSB_ERR_INVALID; }
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaEventSynchronize(e->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->loop_ms += ms;
  if (stats) return ssb_get_stats(e, stats);
  return SSB_OK;
}

int ssb_get_stats(ssb_engine* e, ssb_stats* s) {
  if (!e || !s) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  int rc = e->sync_ctl();
  if (rc) return rc;
  const Ctl& c = *e->h_ctl;
  std::memset(s, 0, sizeof(*s));
  s->rounds = c.round;
  s->processed = (int64_t)c.n_processed;
  s->rollbacks = (int64_t)c.n_rollbacks;
  s->undone = (int64_t)c.n_undone;
  s->antis_sent = (int64_t)c.n_antis;
  s->annihilated = (int64_t)c.n_annihilated;
  s->duplicates = (int64_t)c.n_dups;
  s->committed = (int64_t)c.n_committed;
  s->beyond_finish = (int64_t)c.n_beyond;
  s->remote_sent = (int64_t)c.n_remote;
  s->gvt_time = c.gvt == KEY_MAX ? INT64_MAX : (int64_t)(c.gvt >> 32);
  s->gvt_id = c.gvt == KEY_MAX ? INT64_MAX : (int64_t)(c.gvt & 0xFFFFFFFFull);
  s->pool_high_water = (int64_t)(e->n_chunks - 1 - c.free_low) << e->v.ch_log;
  s->log_dropped = (int64_t)c.log_dropped;
  s->loop_ms = e->loop_ms;
  s->fix_rounds = (int64_t)c.n_fix_rounds;
  s->fix_turns = (int64_t)c.n_fix_total;
  s->exchanges = (int64_t)c.xround;
  return SSB_OK;
}

int ssb_committed_digest(ssb_engine* e, uint64_t d[4]) {
  if (!e || !d) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  int rc = e->sync_ctl();
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) d[i] = e->h_ctl->digest[i];
  return SSB_OK;
}

// a committed log that overflowed is an error for every drain call: the caller would otherwise print an incomplete
// set of result lines with exit status 0
static int log_overflow(ssb_engine* e) {
  if (!e->h_ctl->log_dropped) return SSB_OK;
  e->set_error("committed log overflowed: " + std::to_string(e->h_ctl->log_dropped) +
               " records dropped (raise max_committed or drain more often)");
  return SSB_ERR_CAPACITY;
}

// log fully drained: rewind so that the buffer can be reused by later windows
static int rewind_log(ssb_engine* e, u64 have) {
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  if (e->log_drained == have && have == e->h_ctl->log_count) {
    CK(cudaMemsetAsync(&e->v.ctl->log_count, 0, sizeof(u64), e->stream));
    CK(cudaMemsetAsync(&e->v.ctl->log_stable, 0, sizeof(u64), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->log_drained = 0;
  }
  return SSB_OK;
}

int ssb_drain_committed(ssb_engine* e, ssb_record* out, int64_t cap, int64_t* n) {
  if (!e || !n || (!out && cap) || cap < 0) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  int rc = e->sync_ctl();
  if (rc) return rc;
  if ((rc = log_overflow(e))) return rc;
  u64 have = std::min<u64>(e->h_ctl->log_count, e->v.cap_log);
  u64 avail = have - e->log_drained;
  u64 m = std::min<u64>(avail, (u64)cap);
  *n = (int64_t)m;
  const u64 kChunk = 1u << 22;
  if (m && e->unpack_cap < std::min<u64>(m, kChunk)) {
    if (e->d_unpack) cudaFree(e->d_unpack);
    e->unpack_cap = std::min<u64>(std::max<u64>(m, 1 << 16), kChunk);
    CK(dalloc(&e->d_unpack, e->unpack_cap * 6));
  }
  u64 done = 0;
  while (done < m) {
    u32 k = (u32)std::min<u64>(m - done, e->unpack_cap);
    SSB_LAUNCH(k_unpack_log, e->grid(k, 256), 256, 0, e->stream, (const Ev*)e->v.log, e->log_drained + done, k, e->d_unpack);
    CK(cudaMemcpyAsync(out + done, e->d_unpack, (size_t)k * sizeof(ssb_record), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    done += k;
  }
  e->log_drained += m;
  return rewind_log(e, have);
}

int ssb_drain_committed_raw(ssb_engine* e, ssb_raw_record* out, int64_t cap, int64_t* n) {
  if (!e || !n || (!out && cap) || cap < 0) return SSB_ERR_INVALID;
  static_assert(sizeof(ssb_raw_record) == sizeof(Ev), "raw record layout");
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  int rc = e->sync_ctl();
  if (rc) return rc;
  if ((rc = log_overflow(e))) return rc;
  u64 have = std::min<u64>(e->h_ctl->log_count, e->v.cap_log);
  u64 m = std::min<u64>(have - e->log_drained, (u64)cap);
  *n = (int64_t)m;
  if (m) {
    CK(cudaMemcpyAsync(out, e->v.log + e->log_drained, m * sizeof(Ev), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  e->log_drained += m;
  return rewind_log(e, have);
}

static int ensure_pack_buffers(ssb_engine* e) {
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  const u64 kChunk = 1u << 22;
  if (e->d_pack[0]) return SSB_OK;
  CK(cudaMalloc(&e->d_pack[0], kChunk * 24));
  CK(cudaMalloc(&e->d_pack[1], kChunk * 24));
  CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&e->pack_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->copy_done[i], cudaEventDisableTiming));
  }
  return SSB_OK;
}

int ssb_drain_committed_packed(ssb_engine* e, ssb_packed_record* out, int64_t cap, int64_t* n) {
  if (!e || !n || (!out && cap) || cap < 0) return SSB_ERR_INVALID;
  static_assert(sizeof(ssb_packed_record) == 24, "packed record layout");
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  int rc = e->sync_ctl();
  if (rc) return rc;
  if ((rc = log_overflow(e))) return rc;
  u64 have = std::min<u64>(e->h_ctl->log_count, e->v.cap_log);
  u64 m = std::min<u64>(have - e->log_drained, (u64)cap);
  *n = (int64_t)m;
  // two staging halves: the pack kernel of chunk k+1 runs while chunk k crosses PCIe
  const u64 kChunk = 1u << 22;
  if (m && (rc = ensure_pack_buffers(e))) return rc;
  u64 done = 0;
  for (int k = 0; done < m; ++k) {
    const int b = k & 1;
    const u32 cnt = (u32)std::min<u64>(m - done, kChunk);
    if (k >= 2) CK(cudaStreamWaitEvent(e->stream, e->copy_done[b], 0));      // the half is free again
    SSB_LAUNCH(k_pack_log, e->grid(cnt, 256), 256, 0, e->stream, (const Ev*)e->v.log, e->log_drained + done, cnt, reinterpret_cast<u32*>(e->d_pack[b]));
    CK(cudaEventRecord(e->pack_done[b], e->stream));
    CK(cudaStreamWaitEvent(e->copy_stream, e->pack_done[b], 0));
    CK(cudaMemcpyAsync(out + done, e->d_pack[b], (size_t)cnt * 24, cudaMemcpyDeviceToHost, e->copy_stream));
    CK(cudaEventRecord(e->copy_done[b], e->copy_stream));
    done += cnt;
  }
  if (m) { CK(cudaStreamSynchronize(e->copy_stream)); CK(cudaStreamSynchronize(e->stream)); }
  e->log_drained += m;
  return rewind_log(e, have);
}

int ssb_run_drain_packed(ssb_engine* e, ssb_packed_record* out, int64_t cap, int64_t* n, ssb_stats* stats) {
  if (!e || !n || (!out && cap) || cap < 0) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  int rc = e->model_ready();
  if (rc) return rc;
#ifdef SSB_EMU
  const bool use_graph = false;      // (the interpreted graph loop runs inside ssb_run: nothing to overlap with)
#else
  const bool use_graph = use_graph_path(e);
#endif
  if (!use_graph || e->log_drained != 0) {      // no device-side loop to overlap with: one after the other
    ssb_stats local;
    if ((rc = ssb_run(e, &local))) return rc;
    if (stats) *stats = local;
    if (local.committed > cap) { set_error("ssb_run_drain_packed: output buffer too small"); return SSB_ERR_CAPACITY; }
    return ssb_drain_committed_packed(e, out, cap, n);
  }
#ifndef SSB_EMU
  if (e->graph_dirty && (rc = e->build_graph())) return rc;
  const u64 kChunk = 1u << 22;
  if ((rc = ensure_pack_buffers(e))) return rc;
  if (!e->h_poll) CK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_poll), sizeof(u64), cudaHostAllocDefault));
  CK(cudaEventRecord(e->ev0, e->stream));
  CK(cudaGraphLaunch(e->graph_exec, e->stream));
  CK(cudaEventRecord(e->ev1, e->stream));
  u64 drained = 0;
  bool finished = false;
  int k = 0;
  // pack + copy [drained, upto) on the copy stream, alternating between the two staging halves
  auto drain_to = [&](u64 upto, bool all) -> int {
    while (drained < upto && (all || upto - drained >= (kChunk >> 2))) {
      const u32 cnt = (u32)std::min<u64>(upto - drained, kChunk);
      if ((u64)cap < drained + cnt) { set_error("ssb_run_drain_packed: output buffer too small"); return SSB_ERR_CAPACITY; }
      const int b = k++ & 1;
      SSB_LAUNCH(k_pack_log, e->grid(cnt, 256), 256, 0, e->copy_stream, (const Ev*)e->v.log, drained, cnt, reinterpret_cast<u32*>(e->d_pack[b]));
      CK(cudaMemcpyAsync(out + drained, e->d_pack[b], (size_t)cnt * 24, cudaMemcpyDeviceToHost, e->copy_stream));
      drained += cnt;
    }
    return SSB_OK;
  };
  while (!finished) {
    finished = cudaEventQuery(e->ev1) == cudaSuccess;
    CK(cudaMemcpyAsync(e->h_poll, finished ? &e->v.ctl->log_count : &e->v.ctl->log_stable, sizeof(u64), cudaMemcpyDeviceToHost, e->copy_stream));
    CK(cudaStreamSynchronize(e->copy_stream));      // also waits for the chunks issued so far: the halves are free
    const u64 upto = std::min<u64>(*e->h_poll, e->v.cap_log);
    const u64 before = drained;
    if ((rc = drain_to(upto, finished))) { cudaStreamSynchronize(e->stream); return rc; }
    if (!finished && drained == before) std::this_thread::sleep_for(std::chrono::microseconds(50));   // nothing new yet
  }
  CK(cudaStreamSynchronize(e->copy_stream));
  CK(cudaEventSynchronize(e->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->loop_ms += ms;
  if ((rc = e->sync_ctl())) return rc;
  e->rounds = e->h_ctl->round;
  if ((rc = e->check_device_error())) return rc;
  if (!e->h_ctl->done) { set_error("window loop ended before GVT reached finish_time"); return SSB_ERR_INVALID; }
  if ((rc = log_overflow(e))) return rc;
  *n = (int64_t)drained;
  CK(cudaMemsetAsync(&e->v.ctl->log_count, 0, sizeof(u64), e->stream));      // fully drained: rewind the log
  CK(cudaMemsetAsync(&e->v.ctl->log_stable, 0, sizeof(u64), e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->log_drained = 0;
  if (stats) return ssb_get_stats(e, stats);
#endif
  return SSB_OK;
}

// ssb_run with the committed records streaming out through two pinned staging halves; packed != 0: 24-byte records
// (k_pack_log on the copy stream), else the raw record + its sent-log entry (two plain copies per chunk)
static int stream_run(ssb_engine* e, ssb_packed_sink psink, ssb_history_sink hsink, void* user, ssb_stats* stats) {
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  constexpr u64 kStreamChunk = 1u << 21;
  const bool packed = psink != nullptr;
  if (!packed && !e->v.log_sent) { set_error("history was not kept: create the engine with SSB_FLAG_HISTORY and max_committed > 0"); return SSB_ERR_INVALID; }
  if (!e->v.cap_log) { set_error("streaming drain needs max_committed > 0"); return SSB_ERR_INVALID; }
  int rc = e->model_ready();
  if (rc) return rc;
  if (e->log_drained != 0) { set_error("streaming drain: the committed log is partly drained already"); return SSB_ERR_INVALID; }
  if ((rc = ensure_pack_buffers(e))) return rc;
  for (int i = 0; i < 2; ++i)
    if (!e->h_stage[i]) CK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_stage[i]), kStreamChunk * 48, cudaHostAllocDefault));
  if (!e->h_poll) CK(cudaHostAlloc(reinterpret_cast<void**>(&e->h_poll), sizeof(u64), cudaHostAllocDefault));
  bool overlapped = false;
#ifndef SSB_EMU
  overlapped = use_graph_path(e);
  if (overlapped) {
    if (e->graph_dirty && (rc = e->build_graph())) return rc;
    CK(cudaEventRecord(e->ev0, e->stream));
    CK(cudaGraphLaunch(e->graph_exec, e->stream));
    CK(cudaEventRecord(e->ev1, e->stream));
  }
#endif
  if (!overlapped && (rc = ssb_run(e, nullptr))) return rc;      // no device-side loop to overlap with: run, then stream out
  u64 drained = 0;
  int k = 0, sink_rc = 0;
  struct { u32 cnt; bool live; } pend[2] = {{0, false}, {0, false}};
  auto flush = [&](int b) -> int {
    if (!pend[b].live) return SSB_OK;
    CK(cudaEventSynchronize(e->copy_done[b]));
    pend[b].live = false;
    if (sink_rc) return SSB_OK;
    if (packed) sink_rc = psink(user, reinterpret_cast<const ssb_packed_record*>(e->h_stage[b]), (int64_t)pend[b].cnt);
    else sink_rc = hsink(user, reinterpret_cast<const ssb_raw_record*>(e->h_stage[b]),
                         reinterpret_cast<const ssb_sent_record*>(e->h_stage[b] + kStreamChunk * 32), (int64_t)pend[b].cnt);
    return SSB_OK;
  };
  auto issue = [&](u64 upto, bool all) -> int {
    while (drained < upto && (all || upto - drained >= (kStreamChunk >> 2)) && !sink_rc) {
      const u32 cnt = (u32)std::min<u64>(upto - drained, kStreamChunk);
      const int b = k++ & 1;
      int r = flush(b);      // the half is free once its previous chunk has been handed to the sink
      if (r) return r;
      if (packed) {
        SSB_LAUNCH(k_pack_log, e->grid(cnt, 256), 256, 0, e->copy_stream, (const Ev*)e->v.log, drained, cnt, reinterpret_cast<u32*>(e->d_pack[b]));
        CK(cudaMemcpyAsync(e->h_stage[b], e->d_pack[b], (size_t)cnt * 24, cudaMemcpyDeviceToHost, e->copy_stream));
      } else {
        CK(cudaMemcpyAsync(e->h_stage[b], e->v.log + drained, (size_t)cnt * sizeof(Ev), cudaMemcpyDeviceToHost, e->copy_stream));
        CK(cudaMemcpyAsync(e->h_stage[b] + kStreamChunk * 32, e->v.log_sent + drained, (size_t)cnt * sizeof(uint4), cudaMemcpyDeviceToHost, e->copy_stream));
      }
      CK(cudaEventRecord(e->copy_done[b], e->copy_stream));
      pend[b].cnt = cnt; pend[b].live = true;
      drained += cnt;
    }
    return SSB_OK;
  };
  for (bool finished = false; !finished;) {
    finished = !overlapped || cudaEventQuery(e->ev1) == cudaSuccess;
    CK(cudaMemcpyAsync(e->h_poll, finished ? &e->v.ctl->log_count : &e->v.ctl->log_stable, sizeof(u64), cudaMemcpyDeviceToHost, e->copy_stream));
    CK(cudaStreamSynchronize(e->copy_stream));
    const u64 upto = std::min<u64>(*e->h_poll, e->v.cap_log);
    const u64 before = drained;
    if ((rc = issue(upto, finished))) { cudaStreamSynchronize(e->stream); return rc; }
    if (!finished && drained == before) {
      if ((rc = flush(k & 1)) || (rc = flush((k + 1) & 1))) { cudaStreamSynchronize(e->stream); return rc; }   // idle: hand over what is ready
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
  if ((rc = flush(k & 1)) || (rc = flush((k + 1) & 1))) return rc;      // in issue order
  CK(cudaStreamSynchronize(e->copy_stream));
  if (overlapped) {
    CK(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->loop_ms += ms;
  }
  if ((rc = e->sync_ctl())) return rc;
  e->rounds = e->h_ctl->round;
  if ((rc = e->check_device_error())) return rc;
  if (!e->h_ctl->done) { set_error("window loop ended before GVT reached finish_time"); return SSB_ERR_INVALID; }
  if ((rc = log_overflow(e))) return rc;
  if (sink_rc) { set_error("streaming drain: the sink returned " + std::to_string(sink_rc)); return SSB_ERR_INVALID; }
  CK(cudaMemsetAsync(&e->v.ctl->log_count, 0, sizeof(u64), e->stream));      // fully drained: rewind the log
  CK(cudaMemsetAsync(&e->v.ctl->log_stable, 0, sizeof(u64), e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->log_drained = 0;
  if (stats) return ssb_get_stats(e, stats);
  return SSB_OK;
}

int ssb_run_stream_packed(ssb_engine* e, ssb_packed_sink sink, void* user, ssb_stats* stats) {
  if (!e || !sink) return SSB_ERR_INVALID;
  return stream_run(e, sink, nullptr, user, stats);
}

int ssb_run_stream_history(ssb_engine* e, ssb_history_sink sink, void* user, ssb_stats* stats) {
  if (!e || !sink) return SSB_ERR_INVALID;
  return stream_run(e, nullptr, sink, user, stats);
}

int ssb_drain_history(ssb_engine* e, ssb_raw_record* rec, ssb_sent_record* sent, int64_t cap, int64_t* n) {
  if (!e || !n || ((!rec || !sent) && cap) || cap < 0) return SSB_ERR_INVALID;
  static_assert(sizeof(ssb_sent_record) == sizeof(uint4), "sent record layout");
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  if (!e->v.log_sent) { set_error("history was not kept: create the engine with SSB_FLAG_HISTORY and max_committed > 0"); return SSB_ERR_INVALID; }
  int rc = e->sync_ctl();
  if (rc) return rc;
  if ((rc = log_overflow(e))) return rc;
  u64 have = std::min<u64>(e->h_ctl->log_count, e->v.cap_log);
  u64 m = std::min<u64>(have - e->log_drained, (u64)cap);
  *n = (int64_t)m;
  if (m) {
    CK(cudaMemcpyAsync(rec, e->v.log + e->log_drained, m * sizeof(Ev), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(sent, e->v.log_sent + e->log_drained, m * sizeof(uint4), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  e->log_drained += m;
  return rewind_log(e, have);
}

// --diff_repeat (R13): the recorded events go back into the calendar as PROCESSED events — their recorded sent-log
// entries become their meta records, every bucket counts as consumed, every LP's newest processed time is that of its
// last recorded event. Nothing is executed. What-if events injected afterwards are stragglers at their LPs: the
// ordinary rollback machinery undoes exactly the recorded events they invalidate (those inside the window at once,
// the later ones when the window reaches them: View::inv) — the reference's lazy history load
// (logical_process.hpp:132-153) with the history resident in HBM.
int ssb_load_history(ssb_engine* e, const ssb_raw_record* rec, const ssb_sent_record* sent, int64_t n) {
  if (!e || ((!rec || !sent) && n) || n < 0) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  int rc = e->model_ready();
  if (rc) return rc;
  if ((rc = e->sync_ctl())) return rc;
  if (e->h_ctl->round != 0 || e->h_ctl->load_count != 0) { set_error("ssb_load_history must be the first load after create / reset"); return SSB_ERR_INVALID; }
  if (!e->v.hist_sent && (rc = e->alloc(&e->v.hist_sent, e->v.cap_batch))) return rc;
  if (!e->d_inv && (rc = e->alloc(&e->d_inv, e->v.n_lp_local))) return rc;
  SSB_LAUNCH(k_fill_u64, e->grid(e->v.n_lp_local, 256), 256, 0, e->stream, e->d_inv, KEY_MAX, (size_t)e->v.n_lp_local);
  e->v.replay = 1;
  Ev* stage = reinterpret_cast<Ev*>(e->d_stage);      // the load staging buffer (64 B per entry) holds a chunk of records
  int64_t done = 0;
  while (done < n) {
    u32 m = (u32)std::min<int64_t>(n - done, e->v.cap_batch);
    CK(cudaMemcpyAsync(stage, rec + done, sizeof(Ev) * m, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->v.hist_sent, sent + done, sizeof(uint4) * m, cudaMemcpyHostToDevice, e->stream));
    SSB_LAUNCH(k_hist_stage, e->grid(m, 256), 256, 0, e->stream, e->v, (const Ev*)stage, m);
    e->launch_route(e->v, 3, 0, 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    done += m;
  }
  e->v.replay = 0;
  SSB_LAUNCH(k_after_history, e->grid(e->v.nb, 256), 256, 0, e->stream, e->v);
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaGetLastError());
  e->v.inv = e->d_inv;
  e->graph_dirty = true;
  if ((rc = e->sync_ctl())) return rc;
  return e->check_device_error();
}

int ssb_set_state_from(ssb_engine* e, int64_t lp_id, int64_t time, const uint32_t* words) {
  if (!e || !words) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  if (!e->v.inv) { set_error("ssb_set_state_from needs a loaded history (ssb_load_history first)"); return SSB_ERR_INVALID; }
  if (lp_id < 0 || lp_id >= e->cfg.n_lp || e->owner_of(lp_id) != (u32)e->cfg.rank) { set_error("state change for an LP this rank does not own"); return SSB_ERR_INVALID; }
  if (time < 0 || time >= 0x80000000ll) { set_error("state change time out of range"); return SSB_ERR_RANGE; }
  const u32 l = e->local_of(lp_id);
  const int sw = e->state_words();
  CK(cudaMemcpyAsync(e->v.state + (size_t)l * sw, words, (size_t)sw * 4, cudaMemcpyHostToDevice, e->stream));
  SSB_LAUNCH(k_invalidate_from, 1, 1, 0, e->stream, e->v, l, (u64)time << 32);
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaGetLastError());
  return SSB_OK;
}

int ssb_reset(ssb_engine* e) {
  if (!e) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  return e->reset_sim();
}

int ssb_set_flags(ssb_engine* e, int32_t flags) {
  if (!e) return SSB_ERR_INVALID;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  e->t_collect();
  e->cfg.flags = flags;
  return SSB_OK;
}

int ssb_kernel_times(ssb_engine* e, const char** names, double* total_ms, int64_t* launches, int cap) {
  if (!e || !names || !total_ms || !launches) return 0;
  static const char* kNames[ssb_engine::K_COUNT] = {"k_min", "k_plan", "k_commit", "k_exec", "k_hist_flag", "k_fix_count",
                                                    "k_fix_segs", "k_fix_scatter", "k_fix_sort", "k_fix_turn",
                                                    "k_route(anti)", "k_route", "k_route(recv)", "k_post"};
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  e->t_collect();
  int k = 0;
  for (int i = 0; i < ssb_engine::K_COUNT && k < cap; ++i) {
    names[k] = kNames[i]; total_ms[k] = e->k_ms[i]; launches[k] = e->k_launches[i]; ++k;
  }
  return k;
}

int ssb_comm_unique_id(void* id128) {
  std::string why;
  if (!id128) return SSB_ERR_INVALID;
  if (!g_nccl.load(why)) { g_create_error = why; return SSB_ERR_COMM; }
  Nccl::UniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc); return SSB_ERR_COMM; }
  std::memcpy(id128, &id, 128);
  return SSB_OK;
}

int ssb_comm_init(ssb_engine* e, const void* id128) {
  if (!e || !id128) return SSB_ERR_INVALID;
  if (e->cfg.world <= 1) return SSB_OK;
  std::string why;
  if (!g_nccl.load(why)) { e->set_error(why); return SSB_ERR_COMM; }
  cudaSetDevice(e->cfg.device);
  Nccl::UniqueId id;
  std::memcpy(&id, id128, 128);
  int rc = g_nccl.CommInitRank(&e->comm, e->cfg.world, id, e->cfg.rank);
  if (rc != 0) { e->set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc)); return SSB_ERR_COMM; }
  return SSB_OK;
}

int ssb_set_sync_interval(ssb_engine* e, int32_t windows) {
  if (!e) return SSB_ERR_INVALID;
  if (windows < 1 || windows > 64) { e->set_error("sync interval out of range (1..64)"); return SSB_ERR_INVALID; }
  e->sync_interval = windows;
  e->graph_dirty = true;
  return SSB_OK;
}

int ssb_comm_export(ssb_engine* e, void* handle64) {
  if (!e || !handle64) return SSB_ERR_INVALID;
  if (e->cfg.world <= 1) { std::memset(handle64, 0, 64); return SSB_OK; }
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!e->mbox_block) {
    CK(cudaMalloc(&e->mbox_block, e->mbox_bytes()));
    CK(cudaMemsetAsync(e->mbox_block, 0, e->mbox_cnt_bytes() + e->mbox_gvt_bytes(), e->stream));
    const u32 hdr[4] = {e->v.cap_send, e->v.cap_anti, (u32)e->mbox_cnt_bytes(), (u32)(e->mbox_cnt_bytes() + e->mbox_gvt_bytes())};
    CK(cudaMemcpyAsync(e->mbox_block, hdr, sizeof(hdr), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->mbox_block));
  std::memcpy(handle64, &h, 64);
  return SSB_OK;
}

int ssb_comm_import(ssb_engine* e, const void* handles, int32_t world) {
  if (!e || !handles || world != e->cfg.world) return SSB_ERR_INVALID;
  if (world <= 1) return SSB_OK;
  if (!e->mbox_block) { e->set_error("ssb_comm_export must be called first"); return SSB_ERR_INVALID; }
  cudaSetDevice(e->cfg.device);
  auto set_error = [&](const std::string& s) { e->set_error(s); };
  e->peer_blocks.assign((size_t)world, nullptr);
  for (int r = 0; r < world; ++r) {
    void* base = nullptr;
    if (r == e->cfg.rank) {
      base = e->mbox_block;
    } else {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
      CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
      e->peer_blocks[(size_t)r] = base;
    }
    e->v.peer_base[r] = reinterpret_cast<char*>(base);
    u32 hdr[4];      // the geometry of that partition's mailbox
    CK(cudaMemcpy(hdr, base, sizeof(hdr), cudaMemcpyDeviceToHost));
    e->v.mb_cap_send[r] = hdr[0]; e->v.mb_cap_anti[r] = hdr[1]; e->v.mb_gvt_off[r] = hdr[2]; e->v.mb_data_off[r] = hdr[3];
  }
  e->v.p2p = 1;
  e->graph_dirty = true;
  return SSB_OK;
}

}  // extern "C"
