// fp64_probe.cu — B200 FP64 pipe: dependent-issue latency of DFMA and throughput per SM vs resident warps and ILP.
// (run on the GPU box: build/fp64_probe)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void chain(double* out, int iters, long long* cyc) {
  double a[ILP];
  for (int k = 0; k < ILP; ++k) a[k] = 1.0 + threadIdx.x * 1e-9 + k;
  const double m = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], m, c);
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < ILP; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm, int iters, double* d, long long* dc) {
  const int threads = 128, blocks_per_sm = warps_per_sm / 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = 148 * (blocks_per_sm > 0 ? blocks_per_sm : 1), thr = blocks_per_sm > 0 ? threads : 32 * warps_per_sm;
  chain<ILP><<<grid, thr>>>(d, 100, dc);
  cudaEventRecord(e0); chain<ILP><<<grid, thr>>>(d, iters, dc); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  const double inst = (double)grid * thr / 32 * iters * ILP;
  printf("warps/SM %2d  ILP %d : %.1f cycles per dependent DFMA step (block 0), %.2f warp-DFMA/cycle/SM (=%.1f lanes/clk/SM at %.0f MHz est.)\n",
         warps_per_sm, ILP, (double)c / iters, inst / 148 / (double)c, inst / 148 / (double)c * 32, (double)c / (ms * 1e3));
}
int main() {
  double* d; long long* dc; cudaMalloc(&d, 148 * 32 * 128 * 8 * 4); cudaMalloc(&dc, 8);
  for (int w : {1, 4, 8, 16, 24, 32, 48, 64}) { run<1>(w, 20000, d, dc); run<2>(w, 20000, d, dc); run<4>(w, 10000, d, dc); }
  return 0;
}
