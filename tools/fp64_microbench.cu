// Microbenchmarks that ground DESIGN.md's latency model of the sequential sweeps on B200:
//   DFMA dependent-chain latency, DFMA issue throughput per SMSP (1 warp, ILP 8), FP64 sincos /
//   division / sqrt chain latency, and L2-hit / DRAM dependent-load latency.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double* out, double a, double b, int iters, long long* cyc) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_dfma_ilp(double* out, double a, double b, int iters, long long* cyc) {
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
__global__ void k_fn_chain(double* out, int iters, long long* cyc) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { double s, c; sincos(x, &s, &c); x = s + c; }
    if (OP == 1) x = 1.0 / (x + 1.5);
    if (OP == 2) x = sqrt(x + 2.0);
    if (OP == 3) x = sin(x);
    if (OP == 4) x = atan2(x, 2.7);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_chase(const int* next, int start, int iters, int* sink, long long* cyc) {
  int p = start;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = next[p];
  long long t1 = clock64();
  *sink = p;
  cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps = 1; warps <= 8; warps *= 2) {
    k_dfma_chain<<<1, 32 * warps>>>(out, 0.999, 0.001, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dfma dependent chain, %d warp(s)/SM: %.2f cycles per DFMA (per warp)\n", warps, (double)h / (iters * 16));
  }
  for (int warps = 1; warps <= 16; warps *= 2) {
    k_dfma_ilp<<<1, 32 * warps>>>(out, 0.999, 0.001, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dfma ILP8, %d warp(s)/SM: %.2f cycles per warp-DFMA -> %.1f DFMA lanes/clk/SM\n", warps,
           (double)h / (iters * 16), 32.0 * warps * iters * 16 / (double)h);
  }
  const char* names[] = {"sincos+add", "1/(x+1.5)", "sqrt(x+2)", "sin", "atan2"};
  k_fn_chain<0><<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%s chain: %.1f cycles\n", names[0], (double)h / 1024);
  k_fn_chain<1><<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%s chain: %.1f cycles\n", names[1], (double)h / 1024);
  k_fn_chain<2><<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%s chain: %.1f cycles\n", names[2], (double)h / 1024);
  k_fn_chain<3><<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%s chain: %.1f cycles\n", names[3], (double)h / 1024);
  k_fn_chain<4><<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%s chain: %.1f cycles\n", names[4], (double)h / 1024);
  // pointer chase: 32 MB (L2 resident after warm-up) and 2 GB (DRAM), stride 4 KB + odd offset
  for (int pass = 0; pass < 2; ++pass) {
    size_t n = pass == 0 ? (8u << 20) : (512u << 20);  // ints
    int* next; cudaMalloc(&next, n * 4);
    int* hn = (int*)malloc(n * 4);
    size_t stride = 1031 * 16;  // ints (66 KB) -> different lines/pages
    for (size_t i = 0; i < n; ++i) hn[i] = (int)((i + stride) % n);
    cudaMemcpy(next, hn, n * 4, cudaMemcpyHostToDevice);
    int* sink; cudaMalloc(&sink, 4);
    int it = pass == 0 ? 400 : 2000;
    k_chase<<<1, 1>>>(next, 0, it, sink, cyc);            // warm (L2 for the small one)
    k_chase<<<1, 1>>>(next, pass == 0 ? 0 : 12345, it, sink, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent load latency, %s: %.0f cycles\n", pass == 0 ? "L2-resident 32 MB set" : "2 GB set (DRAM)", (double)h / it);
    cudaFree(next); free(hn);
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock attr %d kHz\n", clk);
  return 0;
}
