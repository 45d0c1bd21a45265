// Device check that altro_b200/csrc/fastmath.cuh reproduces the bits of CUDA's sincos() and 1.0/x
// on its declared range.  nvcc -arch=sm_100a -fmad=false -O3 tools/fastmath_check.cu -o /tmp/fmc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../altro_b200/csrc/fastmath.cuh"
using namespace altro_b200;

__global__ void k_check(const double* xs, int n, unsigned long long* bad, double* ex) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xs[i];
  if (sincos_in_range(x)) {
    double s0, c0, s1, c1;
    sincos(x, &s0, &c0);
    sincos_inrange(x, &s1, &c1);
    if (__double_as_longlong(s0) != __double_as_longlong(s1) || __double_as_longlong(c0) != __double_as_longlong(c1)) {
      if (atomicAdd(bad, 1ull) == 0) { ex[0] = x; ex[1] = s0; ex[2] = s1; ex[3] = c0; ex[4] = c1; }
    }
  }
  if (rcp_in_range(x)) {
    const double r0 = 1.0 / x, r1 = rcp_inrange(x);
    if (__double_as_longlong(r0) != __double_as_longlong(r1)) {
      if (atomicAdd(bad + 1, 1ull) == 0) { ex[5] = x; ex[6] = r0; ex[7] = r1; }
    }
  } else {
    atomicAdd(bad + 2, 1ull);
  }
}

int main() {
  const int n = 1 << 24;
  double* h = (double*)malloc(n * sizeof(double));
  srand48(7);
  for (int i = 0; i < n; ++i) {
    const int kind = i & 7;
    double v;
    if (kind < 3) v = (drand48() * 2 - 1) * 10.0;                    // angles
    else if (kind == 3) v = (drand48() * 2 - 1) * 1e-3;
    else if (kind == 4) v = (drand48() * 2 - 1) * 2.0e9;              // up to the range limit
    else if (kind == 5) v = ldexp(drand48() * 2 - 1, (int)(drand48() * 600) - 300);
    else if (kind == 6) v = (drand48() * 2 - 1) * 1e5;
    else v = ldexp(1.0 + drand48(), (int)(drand48() * 2098) - 1074) * (drand48() < 0.5 ? -1 : 1);  // whole exponent range
    h[i] = v;
  }
  h[0] = 0.0; h[1] = -0.0; h[2] = 1e-310; h[3] = 2147483647.9; h[4] = 3.141592653589793; h[5] = 1.5707963267948966;
  double *d, *ex; unsigned long long* bad;
  cudaMalloc(&d, n * sizeof(double)); cudaMalloc(&bad, 3 * sizeof(unsigned long long)); cudaMalloc(&ex, 8 * sizeof(double));
  cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemset(bad, 0, 3 * sizeof(unsigned long long));
  k_check<<<n / 256, 256>>>(d, n, bad, ex);
  unsigned long long hb[3]; double hex[8];
  cudaMemcpy(hb, bad, sizeof(hb), cudaMemcpyDeviceToHost);
  cudaMemcpy(hex, ex, sizeof(hex), cudaMemcpyDeviceToHost);
  printf("fastmath check over %d inputs: sincos mismatches %llu, rcp mismatches %llu (rcp out of range: %llu) -- %s\n", n, hb[0], hb[1], hb[2],
         cudaGetLastError() == cudaSuccess ? "ok" : cudaGetErrorString(cudaGetLastError()));
  if (hb[0]) printf("  sincos example x=%.17g lib s=%.17g mine s=%.17g lib c=%.17g mine c=%.17g\n", hex[0], hex[1], hex[2], hex[3], hex[4]);
  if (hb[1]) printf("  rcp example x=%.17g lib=%.17g mine=%.17g\n", hex[5], hex[6], hex[7]);
  return (hb[0] || hb[1]) ? 1 : 0;
}
