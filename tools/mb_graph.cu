// micro-benchmark: cost of a CUDA-graph WHILE loop iteration vs. the kernel nodes inside it (B200).
//   nvcc -arch=sm_100a -o mb_graph mb_graph.cu && ./mb_graph
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__global__ void k_tiny(int* p) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1); }
__global__ void k_cond(int* counter, int iters, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { int c = atomicAdd(counter, 1) + 1; cudaGraphSetConditional(h, c < iters ? 1u : 0u); }
}
__global__ void k_if(cudaGraphConditionalHandle h, unsigned v) { if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(h, v); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static int run(int n_nodes, int unroll, int with_if, int grid) {
  cudaStream_t s; CK(cudaStreamCreate(&s));
  int* d; CK(cudaMalloc(&d, 8)); CK(cudaMemset(d, 0, 8));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h, hif;
  CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t nw; CK(cudaGraphAddNode(&nw, g, nullptr, 0, &p));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  const int iters = 2000;
  std::vector<cudaGraphNode_t> deps;
  for (int u = 0; u < unroll; ++u) {
    if (with_if) CK(cudaGraphConditionalHandleCreate(&hif, g, 0, cudaGraphCondAssignDefault));
    CK(cudaStreamBeginCaptureToGraph(s, body, deps.empty() ? nullptr : deps.data(), nullptr, deps.size(), cudaStreamCaptureModeThreadLocal));
    k_cond<<<1, 32, 0, s>>>(d, iters * unroll, h);
    for (int i = 1; i < n_nodes; ++i) k_tiny<<<grid, 256, 0, s>>>(d + 1);
    if (with_if) k_if<<<1, 32, 0, s>>>(hif, 0u);
    cudaStreamCaptureStatus st; const cudaGraphNode_t* dd; size_t nd;
    CK(cudaStreamGetCaptureInfo(s, &st, nullptr, nullptr, &dd, &nd));
    deps.assign(dd, dd + nd);
    cudaGraph_t out; CK(cudaStreamEndCapture(s, &out));
    if (with_if) {
      cudaGraphNodeParams q = {}; q.type = cudaGraphNodeTypeConditional; q.conditional.handle = hif; q.conditional.type = cudaGraphCondTypeIf; q.conditional.size = 1;
      cudaGraphNode_t nif; CK(cudaGraphAddNode(&nif, body, deps.data(), deps.size(), &q));
      cudaGraph_t b2 = q.conditional.phGraph_out[0];
      CK(cudaStreamBeginCaptureToGraph(s, b2, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
      k_tiny<<<1, 32, 0, s>>>(d + 1);
      CK(cudaStreamEndCapture(s, &out));
      deps.assign(1, nif);
    }
  }
  cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(d, 0, 8, s));
    cudaEventRecord(a, s); CK(cudaGraphLaunch(ge, s)); cudaEventRecord(b, s); CK(cudaStreamSynchronize(s));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep == 2) printf("nodes/round %d unroll %d if %d grid %d: %.2f us per round\n", n_nodes, unroll, with_if, grid, 1000.0 * ms / (iters));
  }
  return 0;
}
int main() {
  run(1, 1, 0, 1); run(2, 1, 0, 1); run(4, 1, 0, 1); run(8, 1, 0, 1);
  run(4, 1, 0, 1184); run(4, 4, 0, 1184); run(4, 1, 1, 1184); run(4, 4, 1, 1184); run(1, 8, 0, 1);
  return 0;
}
