// Which HBM layout / load mechanism streams fastest for the sequential sweeps?  One warp = 32
// problems, walks N knots, consumes E doubles per problem per knot with a dependent DFMA chain.
//   layout 0  "problem-fastest SoA": element (k,e) of problem b at  ((k*E+e)*Bp + b)   -> per warp
//             and knot E separate 256-byte pieces, Bp*8 bytes apart
//   layout 1  "knot records (AoSoA)": group g = b/32 owns records [g][k][e][32]        -> per warp
//             and knot ONE contiguous E*256-byte record
// mechanisms: plain LDG | per-lane cp.async (LDGSTS) ring, depth 3 | cp.async.bulk (TMA 1-D) ring
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int E, int LAYOUT>
__global__ void __launch_bounds__(32) k_walk(const double* __restrict__ buf, long Bp, int N, double* out) {
  const long b = (long)blockIdx.x * 32 + threadIdx.x;
  const long g = blockIdx.x;
  double acc = 0.0;
  for (int k = 0; k < N; ++k) {
    double v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const long idx = LAYOUT == 0 ? ((long)(k * E + e) * Bp + b) : ((g * N + k) * E + e) * 32 + threadIdx.x;
      v[e] = buf[idx];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) acc = fma(acc, 0.999, v[e]);
  }
  out[b] = acc;
}

// per-lane cp.async ring
template <int E, int LAYOUT, int D>
__global__ void __launch_bounds__(32) k_walk_cpasync(const double* __restrict__ buf, long Bp, int N, double* out) {
  extern __shared__ double ring[];
  const long b = (long)blockIdx.x * 32 + threadIdx.x;
  const long g = blockIdx.x;
  double* mine = ring + threadIdx.x;
  auto fetch = [&](int k, int s) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const long idx = LAYOUT == 0 ? ((long)(k * E + e) * Bp + b) : ((g * N + k) * E + e) * 32 + threadIdx.x;
      unsigned d = (unsigned)__cvta_generic_to_shared(mine + (s * E + e) * 32);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(buf + idx) : "memory");
    }
  };
  for (int j = 0; j < D; ++j) { if (j < N) fetch(j, j); asm volatile("cp.async.commit_group;" ::: "memory"); }
  double acc = 0.0;
  int s = 0;
  for (int k = 0; k < N; ++k) {
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
    double v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = mine[(s * E + e) * 32];
    if (k + D < N) fetch(k + D, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
    s = (s + 1 == D) ? 0 : s + 1;
#pragma unroll
    for (int e = 0; e < E; ++e) acc = fma(acc, 0.999, v[e]);
  }
  out[b] = acc;
}

// TMA 1-D bulk copy ring (knot records only): one elected lane issues ONE copy per knot
template <int E, int D>
__global__ void __launch_bounds__(32) k_walk_bulk(const double* __restrict__ buf, int N, double* out) {
  extern __shared__ __align__(128) double ring[];
  __shared__ __align__(8) unsigned long long bar[D];
  const long b = (long)blockIdx.x * 32 + threadIdx.x;
  const long g = blockIdx.x;
  const int lane = threadIdx.x;
  constexpr unsigned BYTES = E * 32 * 8;
  if (lane == 0) {
    for (int j = 0; j < D; ++j) {
      unsigned a = (unsigned)__cvta_generic_to_shared(&bar[j]);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto fetch = [&](int k, int s) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&bar[s]);
    unsigned d = (unsigned)__cvta_generic_to_shared(ring + (size_t)s * E * 32);
    const double* src = buf + ((g * N + k) * (long)E) * 32;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(BYTES), "r"(a) : "memory");
  };
  if (lane == 0) for (int j = 0; j < D && j < N; ++j) fetch(j, j);
  double acc = 0.0;
  int s = 0; unsigned phase = 0;
  for (int k = 0; k < N; ++k) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&bar[s]);
    unsigned ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(a), "r"(phase) : "memory");
    }
    double v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ring[((size_t)s * E + e) * 32 + lane];
    __syncwarp();
    if (lane == 0 && k + D < N) fetch(k + D, s);
    s += 1; if (s == D) { s = 0; phase ^= 1; }
#pragma unroll
    for (int e = 0; e < E; ++e) acc = fma(acc, 0.999, v[e]);
  }
  out[b] = acc;
}

template <class F>
float timeit(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return ms / 5;
}

int main() {
  const int N = 100;
  constexpr int E = 54, D = 3;
  const size_t smem = (size_t)D * E * 32 * 8;
  cudaFuncSetAttribute(k_walk_cpasync<E, 0, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_walk_cpasync<E, 1, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_walk_bulk<E, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (long B : {16384L, 65536L, 163840L}) {
    const long Bp = B;
    size_t bytes = (size_t)N * E * Bp * 8;
    double* buf; double* out;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    cudaMalloc(&out, Bp * 8);
    const int W = (int)(B / 32);
    float t[5];
    t[0] = timeit([&] { k_walk<E, 0><<<W, 32>>>(buf, Bp, N, out); });
    t[1] = timeit([&] { k_walk<E, 1><<<W, 32>>>(buf, Bp, N, out); });
    t[2] = timeit([&] { k_walk_cpasync<E, 0, D><<<W, 32, smem>>>(buf, Bp, N, out); });
    t[3] = timeit([&] { k_walk_cpasync<E, 1, D><<<W, 32, smem>>>(buf, Bp, N, out); });
    t[4] = timeit([&] { k_walk_bulk<E, D><<<W, 32, smem>>>(buf, N, out); });
    const char* nm[5] = {"SoA LDG", "records LDG", "SoA cp.async x3", "records cp.async x3", "records TMA-bulk x3"};
    printf("B=%ld E=%d N=%d (%.2f GB, %.1f warps/SM):\n", B, E, N, bytes / 1e9, W / 148.0);
    for (int i = 0; i < 5; ++i) printf("   %-22s %.3f ms = %5.0f GB/s\n", nm[i], t[i], bytes / t[i] / 1e6);
    cudaFree(buf); cudaFree(out);
  }
  return 0;
}
