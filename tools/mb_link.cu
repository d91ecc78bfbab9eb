// Development microbenchmark (not part of the product): what bounds the per-arrival passes (k_link / k_exec shape)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_link tools/mb_link.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;
struct __align__(32) Ev { u64 key; u32 dst, src, send, p0, p1, flags; };
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void ld256(const void* p, u32 (&x)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(x[0]),"=r"(x[1]),"=r"(x[2]),"=r"(x[3]),"=r"(x[4]),"=r"(x[5]),"=r"(x[6]),"=r"(x[7]) : "l"(p));
}

// MODE bits: 1 atomicAdd cnt, 2 atomicExch head, 4 rank0 snapshot (dependent), 8 write link+rank, 16 speculative last load, 32 v8 load
template <int MODE, int U>
__global__ void __launch_bounds__(256) k_link(const Ev* ev, u32 n, uint4* lph, const uint2* last, uint4* link, u32* rank_out, u32* sink) {
  u32 per = (n + gridDim.x - 1) / gridDim.x; per = (per + 31u) & ~31u;
  u32 begin = min(n, blockIdx.x * per), end = min(n, begin + per);
  u32 acc = 0;
  for (u32 base = begin; base < end; base += 256 * U) {
    u64 key[U]; u32 dst[U], fl[U]; bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      u32 i = base + u * 256 + threadIdx.x; ok[u] = i < end;
      if (ok[u]) {
        if (MODE & 32) { u32 x[8]; ld256(ev + i, x); key[u] = ((u64)x[1] << 32) | x[0]; dst[u] = x[2]; fl[u] = x[7]; }
        else { const uint4* q = reinterpret_cast<const uint4*>(ev + i); uint4 a = q[0], b = q[1]; key[u] = ((u64)a.y << 32) | a.x; dst[u] = a.z; fl[u] = b.w; }
      }
    }
    u32 rank[U], prev[U]; uint2 l[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (ok[u]) {
      u32* h = reinterpret_cast<u32*>(lph + dst[u]);
      if (MODE & 16) l[u] = last[dst[u]];
      rank[u] = (MODE & 1) ? atomicAdd(h, 1u) : dst[u];
      prev[u] = (MODE & 2) ? atomicExch(h + 1, base + u) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) if (ok[u]) {
      u32 i = base + u * 256 + threadIdx.x;
      u32* h = reinterpret_cast<u32*>(lph + dst[u]);
      if (MODE & 8) { link[i] = make_uint4((u32)key[u], (u32)(key[u] >> 32), prev[u] | (fl[u] << 31), i); rank_out[i] = rank[u]; }
      else acc += rank[u] + prev[u] + (u32)key[u];
      if ((MODE & 4) && rank[u] == 0) { uint2 x = (MODE & 16) ? l[u] : last[dst[u]]; *reinterpret_cast<uint2*>(h + 2) = x; }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

__global__ void k_flush(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}

template <int MODE, int U>
void run(const char* name, int bps, const Ev* ev, u32 n, uint4* lph, u32 n_lp, uint2* last, uint4* link, u32* rank, u32* sink, uint4* fl, size_t nfl) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9, tot = 0; int reps = 6;
  for (int r = 0; r < reps; ++r) {
    k_flush<<<148 * 8, 256>>>(fl, nfl);
    CK(cudaMemsetAsync(lph, 0, sizeof(uint4) * n_lp));
    cudaEventRecord(a);
    k_link<MODE, U><<<148 * bps, 256>>>(ev, n, lph, last, link, rank, sink);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (r) { best = ms < best ? ms : best; tot += ms; }
  }
  printf("%-46s U=%d bps=%d  best %.1f us  avg %.1f us\n", name, U, bps, best * 1000, tot / (reps - 1) * 1000);
}

int main(int argc, char** argv) {
  u32 n = 1456000, n_lp = 1000000; double hot_frac = argc > 1 ? atof(argv[1]) : 0.38;
  std::vector<Ev> h(n);
  u64 s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (u32 i = 0; i < n; ++i) {
    bool hot = (rnd() % 1000) < hot_frac * 1000;
    h[i].key = ((u64)(100000 + rnd() % 10000) << 32) | (u32)(rnd() % 16000000);
    h[i].dst = hot ? rnd() % 10537 : rnd() % n_lp; h[i].src = 0; h[i].send = 0; h[i].p0 = rnd() % 16; h[i].p1 = 0; h[i].flags = 0;
  }
  Ev* ev; uint4 *lph, *link, *fl; uint2* last; u32 *rank, *sink;
  size_t nfl = (size_t)300 << 20 >> 4;
  CK(cudaMalloc(&ev, sizeof(Ev) * n)); CK(cudaMalloc(&lph, sizeof(uint4) * n_lp)); CK(cudaMalloc(&last, sizeof(uint2) * n_lp));
  CK(cudaMalloc(&link, sizeof(uint4) * n)); CK(cudaMalloc(&rank, 4 * n)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&fl, nfl * 16));
  CK(cudaMemcpy(ev, h.data(), sizeof(Ev) * n, cudaMemcpyHostToDevice)); CK(cudaMemset(last, 0, sizeof(uint2) * n_lp));
  printf("hot fraction %.2f\n", hot_frac);
#define R(M, U, B, NAME) run<M, U>(NAME, B, ev, n, lph, n_lp, last, link, rank, sink, fl, nfl)
  R(0, 1, 8, "load only");
  R(32, 1, 8, "load only v8");
  R(32, 4, 8, "load only v8");
  R(32 | 8, 1, 8, "load + write link");
  R(32 | 1, 1, 8, "load + atomicAdd");
  R(32 | 1 | 2, 1, 8, "load + atomicAdd + exch");
  R(32 | 1 | 2 | 8, 1, 8, "load + 2 atomics + write");
  R(32 | 1 | 2 | 4 | 8, 1, 8, "full (dependent rank0 snapshot)");
  R(32 | 1 | 2 | 4 | 8 | 16, 1, 8, "full (speculative last load)");
  R(32 | 1 | 2 | 4 | 8, 2, 8, "full");
  R(32 | 1 | 2 | 4 | 8, 4, 8, "full");
  R(32 | 1 | 2 | 4 | 8, 4, 4, "full");
  R(32 | 1 | 2 | 4 | 8 | 16, 4, 8, "full spec");
  R(32 | 1 | 2 | 4 | 8 | 16, 4, 4, "full spec");
  R(32 | 1 | 4 | 8, 1, 8, "one atomic (add) + snapshot + write");
  R(32 | 1 | 4 | 8, 4, 8, "one atomic (add) + snapshot + write");
  return 0;
}
