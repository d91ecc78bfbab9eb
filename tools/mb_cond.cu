// micro-benchmark: does a kernel node cost more when the kernel contains a (never executed) cudaGraphSetConditional
// call? WHILE loop of `unroll` rounds, each round one condition kernel + `n` work kernels of `grid` x 256 threads.
//   nvcc -arch=sm_100a -o mb_cond mb_cond.cu && ./mb_cond
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <bool CALL>
__global__ void k_work(int* p, cudaGraphConditionalHandle h, int never) {
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1);
  if (CALL && never && threadIdx.x == 0) cudaGraphSetConditional(h, 1u);
}
__global__ void k_cond(int* counter, int iters, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { int c = atomicAdd(counter, 1) + 1; cudaGraphSetConditional(h, c < iters ? 1u : 0u); }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static int run(int n, int unroll, int call, int grid) {
  cudaStream_t s; CK(cudaStreamCreate(&s));
  int* d; CK(cudaMalloc(&d, 8)); CK(cudaMemset(d, 0, 8));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h;
  CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t nw; CK(cudaGraphAddNode(&nw, g, nullptr, 0, &p));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  const int iters = 2000;
  CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  for (int u = 0; u < unroll; ++u) {
    k_cond<<<1, 32, 0, s>>>(d, iters * unroll, h);
    for (int i = 0; i < n; ++i) {
      if (call) k_work<true><<<grid, 256, 0, s>>>(d + 1, h, 0);
      else k_work<false><<<grid, 256, 0, s>>>(d + 1, h, 0);
    }
  }
  cudaGraph_t out; CK(cudaStreamEndCapture(s, &out));
  cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(d, 0, 8, s));
    cudaEventRecord(a, s); CK(cudaGraphLaunch(ge, s)); cudaEventRecord(b, s); CK(cudaStreamSynchronize(s));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep == 2) printf("work kernels/round %d unroll %d call-inside %d grid %d: %.2f us per round\n", n, unroll, call, grid, 1000.0 * ms / iters);
  }
  return 0;
}
int main() {
  for (int grid : {1, 148, 888}) for (int call = 0; call < 2; ++call) run(4, 8, call, grid);
  return 0;
}
