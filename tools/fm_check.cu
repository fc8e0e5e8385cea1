// fm_check.cu — device-side accuracy check of csrc/coflux_fastmath.cuh against the CUDA math library
// (run on the GPU box: build/fm_check).  Prints the worst error of every function over 2^22 random
// arguments and the raw accuracy of the SFU seeds.  Build: see tools/ab_variants.py (nvcc -arch sm_100a).
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "../climaocean.jl_b200/csrc/coflux_fastmath.cuh"
using namespace coflux;

__device__ double u01(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ULL; uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
  return (double)(z >> 11) * 0x1.0p-53;
}
__device__ double ulps(double a, double ref) {
  if (a == ref) return 0.0;
  int e; frexp(ref, &e);
  return fabs(a - ref) / ldexp(1.0, e - 53);
}
__device__ void amax(double* p, double v) {
  unsigned long long* a = (unsigned long long*)p; unsigned long long old = *a, assumed;
  do { assumed = old; if (__longlong_as_double(assumed) >= v) break; old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
}
__global__ void check(double* out) {
  __shared__ double lgt[256], ext[64];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) lgt[k] = (&COFLUX_LOG_TABLE[0][0])[k];
  if (threadIdx.x < 64) ext[threadIdx.x] = COFLUX_EXP_TABLE[threadIdx.x];
  __syncthreads();
  uint64_t s = 0x1234567ULL + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 7919ULL;
  double w[10] = {0};
  for (int it = 0; it < 64; ++it) {
    const double x = ::exp(-40.0 + 80.0 * u01(s));        // 4e-18 … 2e17
    const double y = ::exp(-40.0 + 80.0 * u01(s));
    const double e = -60.0 + 80.0 * u01(s);
    w[0] = fmax(w[0], ulps(fm::rcp(x), 1.0 / x));
    w[1] = fmax(w[1], ulps(fm::div(y, x), y / x));
    w[2] = fmax(w[2], ulps(fm::sqrt(x), ::sqrt(x)));
    w[3] = fmax(w[3], ulps(fm::cbrt(x), ::cbrt(x)));
    { const double l = ::log(x); w[4] = fmax(w[4], fabs(fm::log(x, lgt) - l) / fmax(1.0, fabs(l)) * 0x1.0p52); }
    w[5] = fmax(w[5], ulps(fm::exp(e, ext), ::exp(e)));
    w[6] = fmax(w[6], fabs(fm::rcp_seed(x) * x - 1.0));
    w[7] = fmax(w[7], fabs(fm::rsqrt_seed(x) * ::sqrt(x) - 1.0));
    { const double xn = 0.5 + 1.5 * u01(s); w[8] = fmax(w[8], fabs(fm::log(xn, lgt) - ::log(xn)) * 0x1.0p52); }
  }
  for (int k = 0; k < 9; ++k) amax(out + k, w[k]);
}
int main() {
  double* d; cudaMalloc(&d, 10 * sizeof(double)); cudaMemset(d, 0, 10 * sizeof(double));
  check<<<512, 128>>>(d);
  double h[10]; cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  printf("fm_check (vs CUDA libm, 4.2M args each): rcp %.2f ulp  div %.2f ulp  sqrt %.2f ulp  cbrt %.2f ulp  log %.2f (2^-52·max(1,|ln|))  "
         "exp %.2f ulp  log[0.5,2] abs %.2f (2^-52)\n", h[0], h[1], h[2], h[3], h[4], h[5], h[8]);
  printf("seed accuracy: rcp.approx.ftz.f64 rel err %.3e (2^%.1f)  rsqrt.approx.ftz.f64 rel err %.3e (2^%.1f)\n", h[6], log2(h[6]), h[7], log2(h[7]));
  return 0;
}
