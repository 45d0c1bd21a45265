// tvlqr.cu -- sections A and B of include/altro_b200.h: the reference's inner C-style TVLQR API
// (src/tvlqr/tvlqr.h:15-33) executed on the GPU, for one problem (host pointer tables, same
// argument lists and return convention) and for a batch (problem-major host arrays).
//
// One thread per LQ problem, problem-fastest device layout (device_problem.h).  Dimensions are
// run-time here (the API takes nx/nu arrays), so blocks live in per-thread local arrays sized by
// a compile-time cap; the fused solver (solver_kernels.cuh) has its own fully unrolled,
// register-resident backward sweep for the compiled-in models.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/altro_b200.h"

namespace {

#define TV_CUDA_OK(expr)                                                               \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      fprintf(stderr, "altro_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), \
              __FILE__, __LINE__);                                                     \
      return ALTRO_B200_ERR_NO_DEVICE;                                                 \
    }                                                                                  \
  } while (0)

struct TvArgs {
  int B, n, m, N;
  long S;  // padded batch stride
  const double *A, *Bm, *f, *Q, *R, *H, *q, *r;
  double reg;
  int is_diag;
  double *K, *d, *P, *p, *dV;
  // optional: the action-value expansion of every knot (outputs of the reference, tvlqr.cpp:123-152)
  double *oQxx, *oQuu, *oQux, *oQx, *oQu;
  int* status;
  // forward
  const double* x0;
  double *x, *u, *y;
};

__device__ __forceinline__ double ld(const double* base, long S, long k, int E, int e) {
  return base[(k * E + e) * S];
}

// tvlqr_BackwardPass, tvlqr.cpp:65-195, operation order as in the reference
template <int NMAX, int MMAX>
__global__ void k_tvlqr_backward(TvArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int n = a.n, m = a.m, N = a.N;
  const long S = a.S;
  double P[NMAX * NMAX], p[NMAX];
  double Am[NMAX * NMAX], Bm[NMAX * MMAX];
  double Qxx[NMAX * NMAX], Quu[MMAX * MMAX], Qux[MMAX * NMAX], Qx[NMAX], Qu[MMAX];
  double T1[NMAX * NMAX], T2[MMAX * NMAX], t[NMAX];
  double K[MMAX * NMAX], d[MMAX], L[MMAX * MMAX];
  double dV0 = 0.0, dV1 = 0.0;
  int result = -1;  // TVLQR_SUCCESS

  // terminal cost-to-go, :82-90
  for (int i = 0; i < n * n; ++i) P[i] = 0.0;
  if (a.is_diag) {
    for (int i = 0; i < n; ++i) P[i + n * i] = ld(a.Q + b, S, N, n, i);
  } else {
    for (int i = 0; i < n * n; ++i) P[i] = ld(a.Q + b, S, N, n * n, i);
  }
  for (int i = 0; i < n; ++i) p[i] = ld(a.q + b, S, N, n, i);
  if (a.P)
    for (int i = 0; i < n * n; ++i) a.P[((long)N * n * n + i) * S + b] = P[i];
  if (a.p)
    for (int i = 0; i < n; ++i) a.p[((long)N * n + i) * S + b] = p[i];

  for (int k = N - 1; k >= 0; --k) {
    for (int i = 0; i < n * n; ++i) Am[i] = ld(a.A + b, S, k, n * n, i);
    for (int i = 0; i < n * m; ++i) Bm[i] = ld(a.Bm + b, S, k, n * m, i);
    for (int i = 0; i < n * n; ++i) Qxx[i] = 0.0;
    for (int i = 0; i < m * m; ++i) Quu[i] = 0.0;
    for (int i = 0; i < m * n; ++i) Qux[i] = 0.0;
    if (a.is_diag) {  // :125-128
      for (int i = 0; i < n; ++i) Qxx[i + n * i] = ld(a.Q + b, S, k, n, i);
      for (int i = 0; i < m; ++i) Quu[i + m * i] = ld(a.R + b, S, k, m, i);
    } else {  // :129-133
      for (int i = 0; i < n * n; ++i) Qxx[i] = ld(a.Q + b, S, k, n * n, i);
      for (int i = 0; i < m * m; ++i) Quu[i] = ld(a.R + b, S, k, m * m, i);
      for (int i = 0; i < m * n; ++i) Qux[i] = ld(a.H + b, S, k, m * n, i);
    }
    // T1 = A' P ; Qxx += T1 A      :135-136
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s = fma(Am[l + n * i], P[l + n * j], s);
        T1[i + n * j] = s;
      }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s = fma(T1[i + n * l], Am[l + n * j], s);
        Qxx[i + n * j] += s;
      }
    // T2 = B' P ; Quu += T2 B ; Qux += T2 A     :139-143
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s = fma(Bm[l + n * i], P[l + n * j], s);
        T2[i + m * j] = s;
      }
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s = fma(T2[i + m * l], Bm[l + n * j], s);
        Quu[i + m * j] += s;
      }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s = fma(T2[i + m * l], Am[l + n * j], s);
        Qux[i + m * j] += s;
      }
    // t = p+ + P+ f ; Qx = q + A' t ; Qu = r + B' t     :147-152
    for (int i = 0; i < n; ++i) {
      double s = p[i];
      for (int l = 0; l < n; ++l) s = fma(P[i + n * l], ld(a.f + b, S, k, n, l), s);
      t[i] = s;
    }
    for (int i = 0; i < n; ++i) {
      double s = ld(a.q + b, S, k, n, i);
      for (int l = 0; l < n; ++l) s = fma(Am[l + n * i], t[l], s);
      Qx[i] = s;
    }
    for (int i = 0; i < m; ++i) {
      double s = ld(a.r + b, S, k, m, i);
      for (int l = 0; l < n; ++l) s = fma(Bm[l + n * i], t[l], s);
      Qu[i] = s;
    }
    if (a.oQxx) {
      for (int i = 0; i < n * n; ++i) a.oQxx[((long)k * n * n + i) * S + b] = Qxx[i];
      for (int i = 0; i < m * m; ++i) a.oQuu[((long)k * m * m + i) * S + b] = Quu[i];
      for (int i = 0; i < m * n; ++i) a.oQux[((long)k * m * n + i) * S + b] = Qux[i];
      for (int i = 0; i < n; ++i) a.oQx[((long)k * n + i) * S + b] = Qx[i];
      for (int i = 0; i < m; ++i) a.oQu[((long)k * m + i) * S + b] = Qu[i];
    }
    // gains, :155-166
    for (int i = 0; i < m * n; ++i) K[i] = Qux[i];
    for (int i = 0; i < m; ++i) d[i] = -Qu[i];
    for (int i = 0; i < m * m; ++i) L[i] = Quu[i];
    for (int i = 0; i < m; ++i) L[i + m * i] += a.reg;
    bool ok = true;
    for (int j = 0; j < m && ok; ++j) {
      double x = L[j + m * j];
      for (int l = 0; l < j; ++l) x = fma(-L[j + m * l], L[j + m * l], x);
      if (x <= 0.0) {
        ok = false;
        break;
      }
      x = sqrt(x);
      L[j + m * j] = x;
      for (int i = j + 1; i < m; ++i) {
        double s = L[i + m * j];
        for (int l = 0; l < j; ++l) s = fma(-L[i + m * l], L[j + m * l], s);
        L[i + m * j] = s / x;
      }
    }
    if (!ok) {  // :162-164: return k; K,d keep the unsolved right-hand sides
      if (a.K)
        for (int i = 0; i < m * n; ++i) a.K[((long)k * m * n + i) * S + b] = K[i];
      if (a.d)
        for (int i = 0; i < m; ++i) a.d[((long)k * m + i) * S + b] = d[i];
      result = k;
      break;
    }
    for (int c = 0; c <= n; ++c) {  // columns of K, then d
      double* xcol = (c < n) ? (K + m * c) : d;
      for (int i = 0; i < m; ++i) {
        double s = xcol[i];
        for (int l = 0; l < i; ++l) s = fma(-L[i + m * l], xcol[l], s);
        xcol[i] = s / L[i + m * i];
      }
      for (int i = m - 1; i >= 0; --i) {
        double s = xcol[i];
        for (int l = i + 1; l < m; ++l) s = fma(-L[l + m * i], xcol[l], s);
        xcol[i] = s / L[i + m * i];
      }
    }
    // cost-to-go, :173-186.  T2 = Quu K ; T1 = K' Qux
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int l = 0; l < m; ++l) s = fma(Quu[i + m * l], K[l + m * j], s);
        T2[i + m * j] = s;
      }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int l = 0; l < m; ++l) s = fma(K[l + m * i], Qux[l + m * j], s);
        T1[i + n * j] = s;
      }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int l = 0; l < m; ++l) s = fma(T2[l + m * i], K[l + m * j], s);
        P[i + n * j] = Qxx[i + n * j] + s;
      }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) P[i + n * j] -= T1[i + n * j];
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) P[i + n * j] -= T1[j + n * i];
    for (int i = 0; i < n; ++i) {
      double s = Qx[i];
      double s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int l = 0; l < m; ++l) {
        s1 = fma(T2[l + m * i], d[l], s1);
        s2 = fma(K[l + m * i], Qu[l], s2);
        s3 = fma(Qux[l + m * i], d[l], s3);
      }
      p[i] = ((s - s1) - s2) + s3;
    }
    // expected decrease, :189-191
    {
      double s0 = 0.0, s1 = 0.0;
      for (int i = 0; i < m; ++i) {
        double qd = 0.0;
        for (int l = 0; l < m; ++l) qd = fma(Quu[i + m * l], d[l], qd);
        s0 = fma(d[i], Qu[i], s0);
        s1 = fma(d[i], qd, s1);
      }
      dV0 += s0;
      dV1 += 0.5 * s1;
    }
    if (a.K)
      for (int i = 0; i < m * n; ++i) a.K[((long)k * m * n + i) * S + b] = K[i];
    if (a.d)
      for (int i = 0; i < m; ++i) a.d[((long)k * m + i) * S + b] = d[i];
    if (a.P)
      for (int i = 0; i < n * n; ++i) a.P[((long)k * n * n + i) * S + b] = P[i];
    if (a.p)
      for (int i = 0; i < n; ++i) a.p[((long)k * n + i) * S + b] = p[i];
  }
  if (a.dV) {
    a.dV[b] = dV0;
    a.dV[S + b] = dV1;
  }
  if (a.status) a.status[b] = result;
}

// tvlqr_ForwardPass, tvlqr.cpp:197-248
template <int NMAX, int MMAX>
__global__ void k_tvlqr_forward(TvArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int n = a.n, m = a.m, N = a.N;
  const long S = a.S;
  double x[NMAX], xn[NMAX], u[MMAX];
  for (int i = 0; i < n; ++i) x[i] = a.x0[(long)i * S + b];
  for (int k = 0; k <= N; ++k) {
    for (int i = 0; i < n; ++i) a.x[((long)k * n + i) * S + b] = x[i];
    if (a.y) {
      for (int i = 0; i < n; ++i) {
        double s = ld(a.p + b, S, k, n, i);
        for (int l = 0; l < n; ++l) s = fma(ld(a.P + b, S, k, n * n, i + n * l), x[l], s);
        a.y[((long)k * n + i) * S + b] = s;
      }
    }
    if (k == N) break;
    for (int i = 0; i < m; ++i) {
      double s = ld(a.d + b, S, k, m, i);
      for (int l = 0; l < n; ++l) s = fma(-ld(a.K + b, S, k, m * n, i + m * l), x[l], s);
      u[i] = s;
      a.u[((long)k * m + i) * S + b] = s;
    }
    for (int i = 0; i < n; ++i) {
      double s = ld(a.f + b, S, k, n, i);
      for (int l = 0; l < n; ++l) s = fma(ld(a.A + b, S, k, n * n, i + n * l), x[l], s);
      for (int l = 0; l < m; ++l) s = fma(ld(a.Bm + b, S, k, n * m, i + n * l), u[l], s);
      xn[i] = s;
    }
    for (int i = 0; i < n; ++i) x[i] = xn[i];
  }
}

// problem-major [B][W] <-> problem-fastest [W][S]
__global__ void k_tv_pm_to_pf(const double* __restrict__ src, int B, long W,
                              double* __restrict__ dst, long ld_) {
  __shared__ double tile[32][33];
  const long j0 = (long)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int bb = b0 + r;
    const long j = j0 + threadIdx.x;
    if (bb < B && j < W) tile[r][threadIdx.x] = src[(long)bb * W + j];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long j = j0 + r;
    const int bb = b0 + threadIdx.x;
    if (bb < B && j < W) dst[j * ld_ + bb] = tile[threadIdx.x][r];
  }
}
__global__ void k_tv_pf_to_pm(const double* __restrict__ src, int B, long W,
                              double* __restrict__ dst, long ld_) {
  __shared__ double tile[32][33];
  const long j0 = (long)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long j = j0 + r;
    const int bb = b0 + threadIdx.x;
    if (bb < B && j < W) tile[r][threadIdx.x] = src[j * ld_ + bb];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int bb = b0 + r;
    const long j = j0 + threadIdx.x;
    if (bb < B && j < W) dst[(long)bb * W + j] = tile[threadIdx.x][r];
  }
}

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  double* alloc(size_t count) {
    void* p = nullptr;
    if (cudaMalloc(&p, (count ? count : 1) * sizeof(double)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return (double*)p;
  }
};

// host [B][W] -> device problem-fastest (new allocation)
int to_device(Scratch& sc, const double* host, int B, long W, long S, double** out) {
  *out = nullptr;
  if (!host) return 0;
  double* stage = sc.alloc((size_t)B * W);
  double* dst = sc.alloc((size_t)W * S);
  if (!stage || !dst) return ALTRO_B200_ERR_NO_DEVICE;
  TV_CUDA_OK(cudaMemcpy(stage, host, (size_t)B * W * 8, cudaMemcpyHostToDevice));
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((B + 31) / 32));
  k_tv_pm_to_pf<<<grid, dim3(32, 8)>>>(stage, B, W, dst, S);
  TV_CUDA_OK(cudaGetLastError());
  *out = dst;
  return 0;
}

int to_host(Scratch& sc, const double* dev, int B, long W, long S, double* host) {
  if (!host) return 0;
  double* stage = sc.alloc((size_t)B * W);
  if (!stage) return ALTRO_B200_ERR_NO_DEVICE;
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((B + 31) / 32));
  k_tv_pf_to_pm<<<grid, dim3(32, 8)>>>(dev, B, W, stage, S);
  TV_CUDA_OK(cudaGetLastError());
  TV_CUDA_OK(cudaMemcpy(host, stage, (size_t)B * W * 8, cudaMemcpyDeviceToHost));
  return 0;
}

template <int NMAX, int MMAX>
void launch_backward(const TvArgs& a) {
  k_tvlqr_backward<NMAX, MMAX><<<(a.B + 31) / 32, 32>>>(a);
}
template <int NMAX, int MMAX>
void launch_forward(const TvArgs& a) {
  k_tvlqr_forward<NMAX, MMAX><<<(a.B + 31) / 32, 32>>>(a);
}

bool have_device() {
  int cnt = 0;
  return cudaGetDeviceCount(&cnt) == cudaSuccess && cnt > 0;
}

// The generic (problem-fastest SoA, one thread per problem) backward pass; Qxx..Qu (host, problem-
// major like the other outputs, may be NULL) receive the action-value expansion of every knot.
static int backward_generic(int batch, int n, int m, int N, const double* A, const double* B,
                            const double* f, const double* Q, const double* R, const double* H,
                            const double* q, const double* r, double reg, bool is_diag, double* K,
                            double* d, double* P, double* p, double* delta_V, int* status,
                            double* Qxx, double* Quu, double* Qux, double* Qx, double* Qu) {
  const long S = ((long)batch + 31) / 32 * 32;
  Scratch sc;
  TvArgs a;
  memset(&a, 0, sizeof(a));
  a.B = batch;
  a.n = n;
  a.m = m;
  a.N = N;
  a.S = S;
  a.reg = reg;
  a.is_diag = is_diag ? 1 : 0;
  double *dA, *dB, *df, *dQ, *dR, *dH = nullptr, *dq, *dr;
  int e = 0;
  const long qw = is_diag ? n : n * n, rw = is_diag ? m : m * m;
  if ((e = to_device(sc, A, batch, (long)N * n * n, S, &dA))) return e;
  if ((e = to_device(sc, B, batch, (long)N * n * m, S, &dB))) return e;
  if ((e = to_device(sc, f, batch, (long)N * n, S, &df))) return e;
  if ((e = to_device(sc, Q, batch, (long)(N + 1) * qw, S, &dQ))) return e;
  if ((e = to_device(sc, R, batch, (long)N * rw, S, &dR))) return e;
  if (!is_diag && (e = to_device(sc, H, batch, (long)N * m * n, S, &dH))) return e;
  if ((e = to_device(sc, q, batch, (long)(N + 1) * n, S, &dq))) return e;
  if ((e = to_device(sc, r, batch, (long)N * m, S, &dr))) return e;
  a.A = dA;
  a.Bm = dB;
  a.f = df;
  a.Q = dQ;
  a.R = dR;
  a.H = dH;
  a.q = dq;
  a.r = dr;
  a.K = sc.alloc((size_t)N * m * n * S);
  a.d = sc.alloc((size_t)N * m * S);
  a.P = sc.alloc((size_t)(N + 1) * n * n * S);
  a.p = sc.alloc((size_t)(N + 1) * n * S);
  a.dV = sc.alloc((size_t)2 * S);
  int* dstatus = (int*)sc.alloc((size_t)S);
  a.status = dstatus;
  if (!a.K || !a.d || !a.P || !a.p || !a.dV || !dstatus) return ALTRO_B200_ERR_NO_DEVICE;
  if (Qxx && Quu && Qux && Qx && Qu) {
    a.oQxx = sc.alloc((size_t)N * n * n * S);
    a.oQuu = sc.alloc((size_t)N * m * m * S);
    a.oQux = sc.alloc((size_t)N * m * n * S);
    a.oQx = sc.alloc((size_t)N * n * S);
    a.oQu = sc.alloc((size_t)N * m * S);
    if (!a.oQxx || !a.oQuu || !a.oQux || !a.oQx || !a.oQu) return ALTRO_B200_ERR_NO_DEVICE;
  }
  if (n <= 4 && m <= 2)
    launch_backward<4, 2>(a);
  else if (n <= 6 && m <= 4)
    launch_backward<6, 4>(a);
  else if (n <= 12 && m <= 4)
    launch_backward<12, 4>(a);
  else
    launch_backward<16, 8>(a);
  TV_CUDA_OK(cudaGetLastError());
  if ((e = to_host(sc, a.K, batch, (long)N * m * n, S, K))) return e;
  if ((e = to_host(sc, a.d, batch, (long)N * m, S, d))) return e;
  if ((e = to_host(sc, a.P, batch, (long)(N + 1) * n * n, S, P))) return e;
  if ((e = to_host(sc, a.p, batch, (long)(N + 1) * n, S, p))) return e;
  if ((e = to_host(sc, a.dV, batch, 2, S, delta_V))) return e;
  if (a.oQxx) {
    if ((e = to_host(sc, a.oQxx, batch, (long)N * n * n, S, Qxx))) return e;
    if ((e = to_host(sc, a.oQuu, batch, (long)N * m * m, S, Quu))) return e;
    if ((e = to_host(sc, a.oQux, batch, (long)N * m * n, S, Qux))) return e;
    if ((e = to_host(sc, a.oQx, batch, (long)N * n, S, Qx))) return e;
    if ((e = to_host(sc, a.oQu, batch, (long)N * m, S, Qu))) return e;
  }
  if (status) TV_CUDA_OK(cudaMemcpy(status, dstatus, sizeof(int) * batch, cudaMemcpyDeviceToHost));
  TV_CUDA_OK(cudaDeviceSynchronize());
  return ALTRO_B200_NO_ERROR;
}


}  // namespace

extern "C" {

// fixed-size, TMA-staged workspace path for the compiled-in shapes (tvlqr_batch.cu); -1 = not compiled in
int altro_b200_tvlqr_cached_backward(int batch, int n, int m, int N, const double* A, const double* B,
                                     const double* f, const double* Q, const double* R, const double* H,
                                     const double* q, const double* r, double reg, bool is_diag, double* K,
                                     double* d, double* P, double* p, double* delta_V, int* status);
int altro_b200_tvlqr_cached_forward(int batch, int n, int m, int N, const double* A, const double* B,
                                    const double* f, const double* K, const double* d, const double* P,
                                    const double* p, const double* x0, double* x, double* u, double* y);

int altro_b200_tvlqr_backward_batch(int batch, int n, int m, int N, const double* A,
                                    const double* B, const double* f, const double* Q,
                                    const double* R, const double* H, const double* q,
                                    const double* r, double reg, bool is_diag, double* K,
                                    double* d, double* P, double* p, double* delta_V,
                                    int* status) {
  if (!A || !B || !f || !Q || !R || !q || !r) return ALTRO_B200_INVALID_POINTER;
  if (!is_diag && !H) return ALTRO_B200_INVALID_POINTER;
  if (n <= 0 || m <= 0 || batch <= 0 || N <= 0) return ALTRO_B200_DIMENSION_UNKNOWN;
  if (n > 16 || m > 8) return ALTRO_B200_ERR_UNSUPPORTED;
  if (!have_device()) return ALTRO_B200_ERR_NO_DEVICE;
  {
    const int e = altro_b200_tvlqr_cached_backward(batch, n, m, N, A, B, f, Q, R, H, q, r, reg, is_diag, K, d,
                                                   P, p, delta_V, status);
    if (e != -1) return e;
  }
  return backward_generic(batch, n, m, N, A, B, f, Q, R, H, q, r, reg, is_diag, K, d, P, p, delta_V, status,
                          nullptr, nullptr, nullptr, nullptr, nullptr);
}

int altro_b200_tvlqr_forward_batch(int batch, int n, int m, int N, const double* A,
                                   const double* B, const double* f, const double* K,
                                   const double* d, const double* P, const double* p,
                                   const double* x0, double* x, double* u, double* y) {
  if (!A || !B || !f || !K || !d || !x0 || !x || !u) return ALTRO_B200_INVALID_POINTER;
  if (y && (!P || !p)) return ALTRO_B200_INVALID_POINTER;
  if (n <= 0 || m <= 0 || batch <= 0 || N <= 0) return ALTRO_B200_DIMENSION_UNKNOWN;
  if (n > 16 || m > 8) return ALTRO_B200_ERR_UNSUPPORTED;
  if (!have_device()) return ALTRO_B200_ERR_NO_DEVICE;
  {
    const int e = altro_b200_tvlqr_cached_forward(batch, n, m, N, A, B, f, K, d, P, p, x0, x, u, y);
    if (e != -1) return e;
  }
  const long S = ((long)batch + 31) / 32 * 32;
  Scratch sc;
  TvArgs a;
  memset(&a, 0, sizeof(a));
  a.B = batch;
  a.n = n;
  a.m = m;
  a.N = N;
  a.S = S;
  double *dA, *dB, *df, *dK, *dd, *dP = nullptr, *dp = nullptr, *dx0;
  int e = 0;
  if ((e = to_device(sc, A, batch, (long)N * n * n, S, &dA))) return e;
  if ((e = to_device(sc, B, batch, (long)N * n * m, S, &dB))) return e;
  if ((e = to_device(sc, f, batch, (long)N * n, S, &df))) return e;
  if ((e = to_device(sc, K, batch, (long)N * m * n, S, &dK))) return e;
  if ((e = to_device(sc, d, batch, (long)N * m, S, &dd))) return e;
  if (y) {
    if ((e = to_device(sc, P, batch, (long)(N + 1) * n * n, S, &dP))) return e;
    if ((e = to_device(sc, p, batch, (long)(N + 1) * n, S, &dp))) return e;
  }
  if ((e = to_device(sc, x0, batch, n, S, &dx0))) return e;
  a.A = dA;
  a.Bm = dB;
  a.f = df;
  a.K = dK;
  a.d = dd;
  a.P = dP;
  a.p = dp;
  a.x0 = dx0;
  a.x = sc.alloc((size_t)(N + 1) * n * S);
  a.u = sc.alloc((size_t)N * m * S);
  a.y = y ? sc.alloc((size_t)(N + 1) * n * S) : nullptr;
  if (!a.x || !a.u || (y && !a.y)) return ALTRO_B200_ERR_NO_DEVICE;
  if (n <= 4 && m <= 2)
    launch_forward<4, 2>(a);
  else if (n <= 6 && m <= 4)
    launch_forward<6, 4>(a);
  else if (n <= 12 && m <= 4)
    launch_forward<12, 4>(a);
  else
    launch_forward<16, 8>(a);
  TV_CUDA_OK(cudaGetLastError());
  if ((e = to_host(sc, a.x, batch, (long)(N + 1) * n, S, x))) return e;
  if ((e = to_host(sc, a.u, batch, (long)N * m, S, u))) return e;
  if (y && (e = to_host(sc, a.y, batch, (long)(N + 1) * n, S, y))) return e;
  TV_CUDA_OK(cudaDeviceSynchronize());
  return ALTRO_B200_NO_ERROR;
}

// ------------------------------------------------------------------ section A (single problem)
// tvlqr.cpp:18-63
int tvlqr_TotalMemSize(const int* nx, const int* nu, int num_horizon, bool is_diag) {
  if (!nx) return 0;
  if (!nu) return 0;
  int mem_size = 0;
  for (int k = 0; k <= num_horizon; ++k) {
    const int n = nx[k];
    mem_size += is_diag ? n : n * n;     // Q
    mem_size += n + n * n + n + n + n;   // q, P, p, x, y
    if (k < num_horizon) {
      const int m = nu[k];
      mem_size += n * n + n * m + n;     // A, B, f
      mem_size += is_diag ? m : m * m;   // R
      mem_size += is_diag ? 0 : m * n;   // H
      mem_size += m;                     // r
      mem_size += m * n + m;             // K, d
      mem_size += 2 * (n * n + m * m + m * n + n + m);  // Qxx..Qu and the _tmp twins
      mem_size += m;                     // u
    }
  }
  mem_size += 2;  // delta_V
  return mem_size * (int)sizeof(lqr_float);
}

static bool uniform_dims(const int* nx, const int* nu, int N, int* n, int* m) {
  if (!nx || !nu || N <= 0) return false;
  *n = nx[0];
  *m = nu[0];
  for (int k = 0; k <= N; ++k)
    if (nx[k] != *n) return false;
  for (int k = 0; k < N; ++k)
    if (nu[k] != *m) return false;
  return true;
}

static void gather(const lqr_float* const* tab, int cnt, int w, std::vector<double>& out) {
  out.resize((size_t)cnt * w);
  for (int k = 0; k < cnt; ++k) memcpy(&out[(size_t)k * w], tab[k], sizeof(double) * w);
}
static void scatter(const std::vector<double>& in, int cnt, int w, lqr_float** tab) {
  if (!tab) return;
  for (int k = 0; k < cnt; ++k)
    if (tab[k]) memcpy(tab[k], &in[(size_t)k * w], sizeof(double) * w);
}

int tvlqr_BackwardPass(const int* nx, const int* nu, int N, const lqr_float* const* A,
                       const lqr_float* const* B, const lqr_float* const* f,
                       const lqr_float* const* Q, const lqr_float* const* R,
                       const lqr_float* const* H, const lqr_float* const* q,
                       const lqr_float* const* r, lqr_float reg, lqr_float** K, lqr_float** d,
                       lqr_float** P, lqr_float** p, lqr_float* delta_V, lqr_float** Qxx,
                       lqr_float** Quu, lqr_float** Qux, lqr_float** Qx, lqr_float** Qu,
                       lqr_float** Qxx_tmp, lqr_float** Quu_tmp, lqr_float** Qux_tmp,
                       lqr_float** Qx_tmp, lqr_float** Qu_tmp, bool linear_only_update,
                       bool is_diag) {
  (void)linear_only_update;  // accepted and ignored, tvlqr.cpp:78
  // the *_tmp tables are the reference's scratch (the GPU keeps scratch in registers); Qxx..Qu are
  // OUTPUTS of the reference (the action-value expansion of every knot, tvlqr.cpp:123-152) and are
  // written when the caller passes all five tables
  (void)Qxx_tmp; (void)Quu_tmp; (void)Qux_tmp; (void)Qx_tmp; (void)Qu_tmp;
  const bool want_q = Qxx && Quu && Qux && Qx && Qu;
  int n, m;
  if (!uniform_dims(nx, nu, N, &n, &m)) return -2;
  std::vector<double> hA, hB, hf, hQ, hR, hH, hq, hr;
  gather(A, N, n * n, hA);
  gather(B, N, n * m, hB);
  gather(f, N, n, hf);
  gather(Q, N + 1, is_diag ? n : n * n, hQ);
  gather(R, N, is_diag ? m : m * m, hR);
  if (!is_diag) gather(H, N, m * n, hH);
  gather(q, N + 1, n, hq);
  gather(r, N, m, hr);
  std::vector<double> hK((size_t)N * m * n), hd((size_t)N * m), hP((size_t)(N + 1) * n * n),
      hp((size_t)(N + 1) * n);
  // pre-load the caller's buffers so knots the pass does not reach keep their old contents
  for (int k = 0; k < N; ++k) {
    if (K && K[k]) memcpy(&hK[(size_t)k * m * n], K[k], sizeof(double) * m * n);
    if (d && d[k]) memcpy(&hd[(size_t)k * m], d[k], sizeof(double) * m);
  }
  double dV[2] = {0, 0};
  int status = -1;
  std::vector<double> hQxx, hQuu, hQux, hQx, hQu;
  int e;
  if (want_q) {
    if (n > 16 || m > 8 || !have_device()) return -2;
    hQxx.resize((size_t)N * n * n);
    hQuu.resize((size_t)N * m * m);
    hQux.resize((size_t)N * m * n);
    hQx.resize((size_t)N * n);
    hQu.resize((size_t)N * m);
    e = backward_generic(1, n, m, N, hA.data(), hB.data(), hf.data(), hQ.data(), hR.data(),
                         is_diag ? nullptr : hH.data(), hq.data(), hr.data(), reg, is_diag, hK.data(),
                         hd.data(), hP.data(), hp.data(), dV, &status, hQxx.data(), hQuu.data(),
                         hQux.data(), hQx.data(), hQu.data());
  } else {
    e = altro_b200_tvlqr_backward_batch(1, n, m, N, hA.data(), hB.data(), hf.data(), hQ.data(),
                                        hR.data(), is_diag ? nullptr : hH.data(), hq.data(),
                                        hr.data(), reg, is_diag, hK.data(), hd.data(), hP.data(),
                                        hp.data(), dV, &status);
  }
  if (e != ALTRO_B200_NO_ERROR) return -2;
  const int k_first = (status == -1) ? 0 : status;  // knots below a failed factorisation are untouched
  if (want_q) {  // the reference has written the expansion of the failing knot as well
    for (int k = N - 1; k >= k_first; --k) {
      if (Qxx[k]) memcpy(Qxx[k], &hQxx[(size_t)k * n * n], sizeof(double) * n * n);
      if (Quu[k]) memcpy(Quu[k], &hQuu[(size_t)k * m * m], sizeof(double) * m * m);
      if (Qux[k]) memcpy(Qux[k], &hQux[(size_t)k * m * n], sizeof(double) * m * n);
      if (Qx[k]) memcpy(Qx[k], &hQx[(size_t)k * n], sizeof(double) * n);
      if (Qu[k]) memcpy(Qu[k], &hQu[(size_t)k * m], sizeof(double) * m);
    }
  }
  for (int k = N; k >= k_first; --k) {
    if (k < N) {
      if (K && K[k]) memcpy(K[k], &hK[(size_t)k * m * n], sizeof(double) * m * n);
      if (d && d[k]) memcpy(d[k], &hd[(size_t)k * m], sizeof(double) * m);
    }
    if (k > k_first || status == -1) {
      if (P && P[k]) memcpy(P[k], &hP[(size_t)k * n * n], sizeof(double) * n * n);
      if (p && p[k]) memcpy(p[k], &hp[(size_t)k * n], sizeof(double) * n);
    }
  }
  if (delta_V) {
    delta_V[0] = dV[0];
    delta_V[1] = dV[1];
  }
  return status;
}

int tvlqr_ForwardPass(const int* nx, const int* nu, int N, const lqr_float* const* A,
                      const lqr_float* const* B, const lqr_float* const* f,
                      const lqr_float* const* K, const lqr_float* const* d,
                      const lqr_float* const* P, const lqr_float* const* p, const lqr_float* x0,
                      lqr_float** x, lqr_float** u, lqr_float** y) {
  int n, m;
  if (!uniform_dims(nx, nu, N, &n, &m)) return -2;
  std::vector<double> hA, hB, hf, hK, hd, hP, hp;
  gather(A, N, n * n, hA);
  gather(B, N, n * m, hB);
  gather(f, N, n, hf);
  gather(K, N, m * n, hK);
  gather(d, N, m, hd);
  if (y) {
    gather(P, N + 1, n * n, hP);
    gather(p, N + 1, n, hp);
  }
  std::vector<double> hx((size_t)(N + 1) * n), hu((size_t)N * m), hy((size_t)(N + 1) * n);
  int e = altro_b200_tvlqr_forward_batch(1, n, m, N, hA.data(), hB.data(), hf.data(), hK.data(),
                                         hd.data(), y ? hP.data() : nullptr,
                                         y ? hp.data() : nullptr, x0, hx.data(), hu.data(),
                                         y ? hy.data() : nullptr);
  if (e != ALTRO_B200_NO_ERROR) return -2;
  scatter(hx, N + 1, n, x);
  scatter(hu, N, m, u);
  if (y) scatter(hy, N + 1, n, y);
  return TVLQR_SUCCESS;
}

}  // extern "C"
