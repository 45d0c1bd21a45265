// tvlqr_batch.cu -- section B of include/altro_b200.h: TVLQR over a BATCH of independent LQ problems
// as a device-resident workspace (src/tvlqr/tvlqr.cpp:65-248 per problem).
//
// The reference's tvlqr_* calls own no memory: the caller hands in per-knot pointer tables
// (tvlqr.h:15-33).  The batched counterpart is a workspace handle that owns ONE record stream in
// HBM, laid out like the solver's knot records (device_problem.h): groups of 32 problems, one
// record per knot, one 256-byte row per block element,
//     element e of field F, knot k, problem b:  rec[(b / 32) * GS + k * R + (rowF + e) * 32 + b % 32]
//     rows:  A n*n | B n*m | f n | Q (n | n*n) | R (m | m*m) | H m*n (dense only) | q n | r m
//            | K m*n | d m | P n*n | p n
// so the backward sweep reads everything it needs at knot k as ONE contiguous range with one TMA
// bulk copy into a shared-memory ring (BulkRing) and writes its results into the same record.
// One warp per group, lane = problem, blocks in registers, (n, m) compile-time for the shapes of
// BASELINE.json's sweep with n <= 6; other shapes use the run-time-dimension kernel of tvlqr.cu.
// No allocation, no synchronisation inside backward(): upload once, run many times.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/altro_b200.h"
#include "linalg.cuh"

using namespace altro_b200;

namespace {

#define TVB_OK(expr)                                                                    \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      fprintf(stderr, "altro_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), \
              __FILE__, __LINE__);                                                      \
      return ALTRO_B200_ERR_NO_DEVICE;                                                  \
    }                                                                                   \
  } while (0)

template <int n, int m, bool DIAG>
struct TvRows {
  static constexpr int qw = DIAG ? n : n * n, rw = DIAG ? m : m * m, hw = DIAG ? 0 : m * n;
  static constexpr int rA = 0, rB = rA + n * n, rF = rB + n * m, rQ = rF + n, rR = rQ + qw, rH = rR + rw,
                       rq = rH + hw, rr = rq + n, kIn = rr + m;
  static constexpr int rK = kIn, rD = rK + m * n, rP = rD + m, rp = rP + n * n, kRows = rp + n;
};

extern __shared__ __align__(128) unsigned char tv_smem[];

// tvlqr_BackwardPass (tvlqr.cpp:65-195), operation order of the reference; one warp per group
template <int n, int m, bool DIAG>
__global__ void __launch_bounds__(32) k_tvlqr_backward_rec(double* rec, long R, long GS, int B, int N, double reg,
                                                            int depth, double* dV, int* status) {
  using L = TvRows<n, m, DIAG>;
  const int g = blockIdx.x, lane = threadIdx.x;
  const int b = g * 32 + lane;
  const bool valid = b < B;
  double* base = rec + (long)g * GS;
  BulkRing ring;
  ring.init(tv_smem, depth, L::kIn * 32, lane == 0);
  __syncwarp();
  auto fetch = [&](int k, int st) {
    ring.expect(st, L::kIn * 256);
    ring.copy(st, 0, base + (long)k * R, L::kIn * 256);
  };
  if (lane == 0)
    for (int j = 0; j < depth; ++j)
      if (N - 1 - j >= 0) fetch(N - 1 - j, j);
  double Pn[n * n], pn[n];
  {  // terminal cost-to-go, :82-90
    const double* t = base + (long)N * R + lane;
#pragma unroll
    for (int i = 0; i < n * n; ++i) Pn[i] = 0.0;
    if constexpr (DIAG) {
#pragma unroll
      for (int i = 0; i < n; ++i) Pn[i + n * i] = t[(L::rQ + i) * 32];
    } else {
#pragma unroll
      for (int i = 0; i < n * n; ++i) Pn[i] = t[(L::rQ + i) * 32];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) pn[i] = t[(L::rq + i) * 32];
    double* o = base + (long)N * R + lane;
#pragma unroll
    for (int i = 0; i < n * n; ++i) o[(L::rP + i) * 32] = Pn[i];
#pragma unroll
    for (int i = 0; i < n; ++i) o[(L::rp + i) * 32] = pn[i];
  }
  double dV0 = 0.0, dV1 = 0.0;
  int result = -1;  // TVLQR_SUCCESS
  bool alive = true;
  for (int k = N - 1; k >= 0; --k) {
    const double* st = ring.wait();
    if (alive) {
      double A[n * n], Bm[n * m], f[n], Qxx[n * n], Quu[m * m], Qux[m * n], Qx[n], Qu[m];
      unstage_block<n * n>(st, L::rA, lane, A);
      unstage_block<n * m>(st, L::rB, lane, Bm);
      unstage_block<n>(st, L::rF, lane, f);
      if constexpr (DIAG) {  // :125-128
#pragma unroll
        for (int i = 0; i < n * n; ++i) Qxx[i] = 0.0;
#pragma unroll
        for (int i = 0; i < m * m; ++i) Quu[i] = 0.0;
#pragma unroll
        for (int i = 0; i < m * n; ++i) Qux[i] = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i) Qxx[i + n * i] = st[(L::rQ + i) * 32 + lane];
#pragma unroll
        for (int i = 0; i < m; ++i) Quu[i + m * i] = st[(L::rR + i) * 32 + lane];
      } else {  // :129-133
        unstage_block<n * n>(st, L::rQ, lane, Qxx);
        unstage_block<m * m>(st, L::rR, lane, Quu);
        unstage_block<L::hw>(st, L::rH, lane, Qux);
      }
      unstage_block<n>(st, L::rq, lane, Qx);
      unstage_block<m>(st, L::rr, lane, Qu);
      {
        double T1[n * n];
        mm<n, n, n, true, false, 0>(A, Pn, T1);    // A' P+            :135
        mm<n, n, n, false, false, 1>(T1, A, Qxx);  // Qxx += (A'P+) A  :136
      }
      {
        double T2[m * n];
        mm<m, n, n, true, false, 0>(Bm, Pn, T2);    // B' P+            :139
        mm<m, m, n, false, false, 1>(T2, Bm, Quu);  // Quu += (B'P+) B  :140
        mm<m, n, n, false, false, 1>(T2, A, Qux);   // Qux += (B'P+) A  :143
      }
      {
        double t[n];  // t = p+ + P+ f                                  :147-148
#pragma unroll
        for (int i = 0; i < n; ++i) t[i] = pn[i];
        mm<n, 1, n, false, false, 1>(Pn, f, t);
        mm<n, 1, n, true, false, 1>(A, t, Qx);   // Qx = q + A' t       :149-150
        mm<m, 1, n, true, false, 1>(Bm, t, Qu);  // Qu = r + B' t       :151-152
      }
      double K[m * n], d[m], Lc[m * m];
#pragma unroll
      for (int i = 0; i < m * n; ++i) K[i] = Qux[i];
#pragma unroll
      for (int i = 0; i < m; ++i) d[i] = -Qu[i];
#pragma unroll
      for (int i = 0; i < m * m; ++i) Lc[i] = Quu[i];
#pragma unroll
      for (int i = 0; i < m; ++i) Lc[i + m * i] += reg;  // Quu + reg I   :157-161
      double* o = base + (long)k * R + lane;
      if (!cholesky<m>(Lc)) {  // :162-164: return the failing knot
        result = k;
        alive = false;
      } else {
        cholesky_solve<m, n>(Lc, K);  // :165
        cholesky_solve<m, 1>(Lc, d);  // :166
        if (valid) {
#pragma unroll
          for (int i = 0; i < m * n; ++i) o[(L::rK + i) * 32] = K[i];
#pragma unroll
          for (int i = 0; i < m; ++i) o[(L::rD + i) * 32] = d[i];
        }
        double QuuK[m * n], KtQux[n * n];  // cost-to-go, :173-186
        mm<m, n, m, false, false, 0>(Quu, K, QuuK);
        mm<n, n, m, true, false, 0>(K, Qux, KtQux);
#pragma unroll
        for (int i = 0; i < n * n; ++i) Pn[i] = Qxx[i];
        mm<n, n, m, true, false, 1>(QuuK, K, Pn);
#pragma unroll
        for (int c = 0; c < n; ++c)
#pragma unroll
          for (int rr_ = 0; rr_ < n; ++rr_) Pn[rr_ + n * c] -= KtQux[rr_ + n * c];
#pragma unroll
        for (int c = 0; c < n; ++c)
#pragma unroll
          for (int rr_ = 0; rr_ < n; ++rr_) Pn[rr_ + n * c] -= KtQux[c + n * rr_];
#pragma unroll
        for (int i = 0; i < n; ++i) pn[i] = Qx[i];
        mm<n, 1, m, true, false, -1>(QuuK, d, pn);
        mm<n, 1, m, true, false, -1>(K, Qu, pn);
        mm<n, 1, m, true, false, 1>(Qux, d, pn);
        if (valid) {
#pragma unroll
          for (int i = 0; i < n * n; ++i) o[(L::rP + i) * 32] = Pn[i];
#pragma unroll
          for (int i = 0; i < n; ++i) o[(L::rp + i) * 32] = pn[i];
        }
        dV0 += dot<m>(d, Qu);  // :189-191
        {
          double Qd_[m];
          mm<m, 1, m, false, false, 0>(Quu, d, Qd_);
          dV1 += 0.5 * dot<m>(d, Qd_);
        }
      }
    }
    __syncwarp();
    if (lane == 0 && k - depth >= 0) fetch(k - depth, ring.s);
    ring.advance();
  }
  if (valid) {
    if (dV) {
      dV[2 * b] = dV0;
      dV[2 * b + 1] = dV1;
    }
    if (status) status[b] = result;
  }
}

// tvlqr_ForwardPass (tvlqr.cpp:197-248): u = d - K x, x+ = f + A x + B u, y = P x + p
template <int n, int m, bool DIAG>
__global__ void __launch_bounds__(32) k_tvlqr_forward_rec(const double* rec, long R, long GS, int B, int N,
                                                           const double* x0, double* X, double* U, double* Y) {
  using L = TvRows<n, m, DIAG>;
  const int g = blockIdx.x, lane = threadIdx.x;
  const int b = g * 32 + lane;
  if (b >= B) return;
  const double* base = rec + (long)g * GS + lane;
  double x[n];
#pragma unroll
  for (int i = 0; i < n; ++i) x[i] = x0[(long)b * n + i];
  for (int k = 0; k <= N; ++k) {
    const double* t = base + (long)k * R;
#pragma unroll
    for (int i = 0; i < n; ++i) X[((long)b * (N + 1) + k) * n + i] = x[i];
    if (Y) {
      double Pk[n * n], y[n];
      load_block<n * n>(t + L::rP * 32, 0, 0, Pk);
      load_block<n>(t + L::rp * 32, 0, 0, y);
      mm<n, 1, n, false, false, 1>(Pk, x, y);
#pragma unroll
      for (int i = 0; i < n; ++i) Y[((long)b * (N + 1) + k) * n + i] = y[i];
    }
    if (k == N) break;
    double K[m * n], u[m], A[n * n], Bm[n * m], xn[n];
    load_block<m * n>(t + L::rK * 32, 0, 0, K);
    load_block<m>(t + L::rD * 32, 0, 0, u);
    mm<m, 1, n, false, false, -1>(K, x, u);  // u = d - K x            :223-224
#pragma unroll
    for (int i = 0; i < m; ++i) U[((long)b * N + k) * m + i] = u[i];
    load_block<n * n>(t + L::rA * 32, 0, 0, A);
    load_block<n * m>(t + L::rB * 32, 0, 0, Bm);
    load_block<n>(t + L::rF * 32, 0, 0, xn);
    mm<n, 1, n, false, false, 1>(A, x, xn);   // x+ = f + A x + B u     :226-228
    mm<n, 1, m, false, false, 1>(Bm, u, xn);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = xn[i];
  }
}

// host problem-major [B][knots][E] <-> rows [row0, row0 + E) of the records of knots [0, knots)
__global__ void k_tv_scatter(const double* __restrict__ src, int B, int knots, int E, int row0, long R, long GS,
                             double* __restrict__ rec) {
  __shared__ double tile[32][33];
  const long W = (long)knots * E;
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
    if (bb < B && j < W) {
      const long k = j / E;
      const int e = (int)(j - k * E);
      rec[(long)(bb >> 5) * GS + k * R + (long)(row0 + e) * 32 + (bb & 31)] = tile[threadIdx.x][r];
    }
  }
}
__global__ void k_tv_gather(const double* __restrict__ rec, int B, int knots, int E, int row0, long R, long GS,
                            double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const long W = (long)knots * E;
  const long j0 = (long)blockIdx.x * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long j = j0 + r;
    const int bb = b0 + threadIdx.x;
    if (bb < B && j < W) {
      const long k = j / E;
      const int e = (int)(j - k * E);
      tile[r][threadIdx.x] = rec[(long)(bb >> 5) * GS + k * R + (long)(row0 + e) * 32 + (bb & 31)];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int bb = b0 + r;
    const long j = j0 + threadIdx.x;
    if (bb < B && j < W) dst[(long)bb * W + j] = tile[threadIdx.x][r];
  }
}

struct RowMap {
  int rA, rB, rF, rQ, rR, rH, rq, rr, kIn, rK, rD, rP, rp, kRows, qw, rw;
};
template <int n, int m, bool DIAG>
RowMap make_rows() {
  using L = TvRows<n, m, DIAG>;
  return RowMap{L::rA, L::rB, L::rF, L::rQ, L::rR, L::rH, L::rq, L::rr, L::kIn, L::rK, L::rD, L::rP, L::rp,
                L::kRows, L::qw, L::rw};
}

typedef void (*bw_fn)(double*, long, long, int, int, double, int, double*, int*, int, size_t, cudaStream_t);
typedef void (*fw_fn)(const double*, long, long, int, int, const double*, double*, double*, double*, int, cudaStream_t);

template <int n, int m, bool DIAG>
void run_bw(double* rec, long R, long GS, int B, int N, double reg, int depth, double* dV, int* status, int G,
            size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(k_tvlqr_backward_rec<n, m, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_tvlqr_backward_rec<n, m, DIAG><<<G, 32, smem, st>>>(rec, R, GS, B, N, reg, depth, dV, status);
}
template <int n, int m, bool DIAG>
void run_fw(const double* rec, long R, long GS, int B, int N, const double* x0, double* X, double* U, double* Y,
            int G, cudaStream_t st) {
  k_tvlqr_forward_rec<n, m, DIAG><<<G, 32, 0, st>>>(rec, R, GS, B, N, x0, X, U, Y);
}

struct Shape {
  int n, m;
  RowMap rows[2];  // [is_diag]
  bw_fn bw[2];
  fw_fn fw[2];
};
#define TV_SHAPE(N_, M_)                                                                         \
  Shape {                                                                                        \
    N_, M_, {make_rows<N_, M_, false>(), make_rows<N_, M_, true>()},                             \
        {run_bw<N_, M_, false>, run_bw<N_, M_, true>}, { run_fw<N_, M_, false>, run_fw<N_, M_, true> } \
  }
const Shape* find_shape(int n, int m) {
  static const Shape shapes[] = {TV_SHAPE(2, 1), TV_SHAPE(4, 2), TV_SHAPE(4, 4), TV_SHAPE(5, 2),
                                 TV_SHAPE(6, 2), TV_SHAPE(6, 3), TV_SHAPE(6, 4)};
  for (const Shape& s : shapes)
    if (s.n == n && s.m == m) return &s;
  return nullptr;
}

}  // namespace

struct altro_b200_tvlqr_ws {
  int B, n, m, N, is_diag, device, G;
  const Shape* shape;
  RowMap rows;
  long R, GS;
  double* rec = nullptr;
  double* dV = nullptr;
  int* status = nullptr;
  double* stage = nullptr;  // transposition staging, sized for the largest field
  long stage_count = 0;
  double *x0 = nullptr, *X = nullptr, *U = nullptr, *Y = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int depth = 2;
  size_t smem = 0;
};

extern "C" {

altro_b200_tvlqr_ws* altro_b200_tvlqr_ws_create(int batch, int n, int m, int N, bool is_diag, int device) {
  if (batch <= 0 || N <= 0) return nullptr;
  const Shape* sh = find_shape(n, m);
  if (!sh) return nullptr;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0 || device >= cnt) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  altro_b200_tvlqr_ws* w = new altro_b200_tvlqr_ws();
  w->B = batch;
  w->n = n;
  w->m = m;
  w->N = N;
  w->is_diag = is_diag ? 1 : 0;
  w->device = device;
  w->G = (batch + 31) / 32;
  w->shape = sh;
  w->rows = sh->rows[w->is_diag];
  w->R = (long)w->rows.kRows * 32;
  w->GS = (long)(N + 1) * w->R;
  const long big = (long)batch * (N + 1) * n * n;
  w->stage_count = big;
  bool ok = cudaMalloc((void**)&w->rec, (size_t)w->G * w->GS * 8) == cudaSuccess &&
            cudaMemset(w->rec, 0, (size_t)w->G * w->GS * 8) == cudaSuccess &&
            cudaMalloc((void**)&w->dV, (size_t)w->G * 32 * 2 * 8) == cudaSuccess &&
            cudaMalloc((void**)&w->status, (size_t)w->G * 32 * 4) == cudaSuccess &&
            cudaMalloc((void**)&w->stage, (size_t)big * 8) == cudaSuccess &&
            cudaMalloc((void**)&w->x0, (size_t)batch * n * 8) == cudaSuccess &&
            cudaMalloc((void**)&w->X, (size_t)batch * (N + 1) * n * 8) == cudaSuccess &&
            cudaMalloc((void**)&w->U, (size_t)batch * N * m * 8) == cudaSuccess &&
            cudaMalloc((void**)&w->Y, (size_t)batch * (N + 1) * n * 8) == cudaSuccess &&
            cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&w->ev0) == cudaSuccess && cudaEventCreate(&w->ev1) == cudaSuccess;
  if (!ok) {
    altro_b200_tvlqr_ws_destroy(w);
    return nullptr;
  }
  // ring depth: as deep as shared memory allows with every group resident (<= kMaxStageDepth)
  int sms = 148, smem_sm = 228 * 1024;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  const int per_sm = (w->G + sms - 1) / sms;
  const size_t stage_bytes = (size_t)w->rows.kIn * 256;
  const size_t budget = (size_t)smem_sm / (per_sm > 0 ? per_sm : 1) - 1024 - 128;
  size_t d = budget / stage_bytes;
  w->depth = (int)(d < 2 ? 2 : (d > (size_t)kMaxStageDepth ? (size_t)kMaxStageDepth : d));
  w->smem = BulkRing::bytes(w->depth, w->rows.kIn * 32);
  return w;
}

void altro_b200_tvlqr_ws_destroy(altro_b200_tvlqr_ws* w) {
  if (!w) return;
  cudaSetDevice(w->device);
  if (w->stream) cudaStreamSynchronize(w->stream);
  for (void* p : {(void*)w->rec, (void*)w->dV, (void*)w->status, (void*)w->stage, (void*)w->x0, (void*)w->X,
                  (void*)w->U, (void*)w->Y})
    if (p) cudaFree(p);
  if (w->ev0) cudaEventDestroy(w->ev0);
  if (w->ev1) cudaEventDestroy(w->ev1);
  if (w->stream) cudaStreamDestroy(w->stream);
  delete w;
}

static int tv_put(altro_b200_tvlqr_ws* w, const double* host, int knots, int E, int row0) {
  if (!host || E == 0) return 0;
  const long W = (long)knots * E;
  TVB_OK(cudaMemcpyAsync(w->stage, host, (size_t)w->B * W * 8, cudaMemcpyHostToDevice, w->stream));
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((w->B + 31) / 32));
  k_tv_scatter<<<grid, dim3(32, 8), 0, w->stream>>>(w->stage, w->B, knots, E, row0, w->R, w->GS, w->rec);
  TVB_OK(cudaGetLastError());
  TVB_OK(cudaStreamSynchronize(w->stream));  // the staging buffer is reused by the next field
  return 0;
}
static int tv_get(altro_b200_tvlqr_ws* w, double* host, int knots, int E, int row0) {
  if (!host) return 0;
  const long W = (long)knots * E;
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((w->B + 31) / 32));
  k_tv_gather<<<grid, dim3(32, 8), 0, w->stream>>>(w->rec, w->B, knots, E, row0, w->R, w->GS, w->stage);
  TVB_OK(cudaGetLastError());
  TVB_OK(cudaMemcpyAsync(host, w->stage, (size_t)w->B * W * 8, cudaMemcpyDeviceToHost, w->stream));
  TVB_OK(cudaStreamSynchronize(w->stream));
  return 0;
}

int altro_b200_tvlqr_ws_upload(altro_b200_tvlqr_ws* w, const double* A, const double* B, const double* f,
                               const double* Q, const double* R, const double* H, const double* q,
                               const double* r) {
  if (!w) return ALTRO_B200_INVALID_POINTER;
  TVB_OK(cudaSetDevice(w->device));
  const RowMap& L = w->rows;
  const int N = w->N, n = w->n, m = w->m;
  int e = 0;
  if ((e = tv_put(w, A, N, n * n, L.rA))) return e;
  if ((e = tv_put(w, B, N, n * m, L.rB))) return e;
  if ((e = tv_put(w, f, N, n, L.rF))) return e;
  if ((e = tv_put(w, Q, N + 1, L.qw, L.rQ))) return e;
  if ((e = tv_put(w, R, N, L.rw, L.rR))) return e;
  if (!w->is_diag && H && (e = tv_put(w, H, N, m * n, L.rH))) return e;
  if ((e = tv_put(w, q, N + 1, n, L.rq))) return e;
  if ((e = tv_put(w, r, N, m, L.rr))) return e;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_tvlqr_ws_backward(altro_b200_tvlqr_ws* w, double reg) {
  if (!w) return ALTRO_B200_INVALID_POINTER;
  TVB_OK(cudaSetDevice(w->device));
  w->shape->bw[w->is_diag](w->rec, w->R, w->GS, w->B, w->N, reg, w->depth, w->dV, w->status, w->G, w->smem,
                           w->stream);
  TVB_OK(cudaGetLastError());
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_tvlqr_ws_time_backward(altro_b200_tvlqr_ws* w, double reg, int reps, float* ms_per_launch) {
  if (!w || !ms_per_launch || reps <= 0) return ALTRO_B200_INVALID_POINTER;
  TVB_OK(cudaSetDevice(w->device));
  int e = altro_b200_tvlqr_ws_backward(w, reg);  // warm-up
  if (e) return e;
  TVB_OK(cudaEventRecord(w->ev0, w->stream));
  for (int i = 0; i < reps; ++i)
    if ((e = altro_b200_tvlqr_ws_backward(w, reg))) return e;
  TVB_OK(cudaEventRecord(w->ev1, w->stream));
  TVB_OK(cudaEventSynchronize(w->ev1));
  float ms = 0.f;
  TVB_OK(cudaEventElapsedTime(&ms, w->ev0, w->ev1));
  *ms_per_launch = ms / reps;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_tvlqr_ws_download(altro_b200_tvlqr_ws* w, double* K, double* d, double* P, double* p,
                                 double* delta_V, int* status) {
  if (!w) return ALTRO_B200_INVALID_POINTER;
  TVB_OK(cudaSetDevice(w->device));
  const RowMap& L = w->rows;
  const int N = w->N, n = w->n, m = w->m;
  int e = 0;
  if ((e = tv_get(w, K, N, m * n, L.rK))) return e;
  if ((e = tv_get(w, d, N, m, L.rD))) return e;
  if ((e = tv_get(w, P, N + 1, n * n, L.rP))) return e;
  if ((e = tv_get(w, p, N + 1, n, L.rp))) return e;
  if (delta_V)
    TVB_OK(cudaMemcpyAsync(delta_V, w->dV, sizeof(double) * 2 * w->B, cudaMemcpyDeviceToHost, w->stream));
  if (status) TVB_OK(cudaMemcpyAsync(status, w->status, sizeof(int) * w->B, cudaMemcpyDeviceToHost, w->stream));
  TVB_OK(cudaStreamSynchronize(w->stream));
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_tvlqr_ws_set_gains(altro_b200_tvlqr_ws* w, const double* K, const double* d, const double* P,
                                  const double* p) {
  if (!w) return ALTRO_B200_INVALID_POINTER;
  TVB_OK(cudaSetDevice(w->device));
  const RowMap& L = w->rows;
  int e = 0;
  if ((e = tv_put(w, K, w->N, w->m * w->n, L.rK))) return e;
  if ((e = tv_put(w, d, w->N, w->m, L.rD))) return e;
  if ((e = tv_put(w, P, w->N + 1, w->n * w->n, L.rP))) return e;
  if ((e = tv_put(w, p, w->N + 1, w->n, L.rp))) return e;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_tvlqr_ws_forward(altro_b200_tvlqr_ws* w, const double* x0, double* x, double* u, double* y) {
  if (!w || !x0 || !x || !u) return ALTRO_B200_INVALID_POINTER;
  TVB_OK(cudaSetDevice(w->device));
  const int N = w->N, n = w->n, m = w->m;
  TVB_OK(cudaMemcpyAsync(w->x0, x0, sizeof(double) * (size_t)w->B * n, cudaMemcpyHostToDevice, w->stream));
  w->shape->fw[w->is_diag](w->rec, w->R, w->GS, w->B, N, w->x0, w->X, w->U, y ? w->Y : nullptr, w->G, w->stream);
  TVB_OK(cudaGetLastError());
  TVB_OK(cudaMemcpyAsync(x, w->X, sizeof(double) * (size_t)w->B * (N + 1) * n, cudaMemcpyDeviceToHost, w->stream));
  TVB_OK(cudaMemcpyAsync(u, w->U, sizeof(double) * (size_t)w->B * N * m, cudaMemcpyDeviceToHost, w->stream));
  if (y) TVB_OK(cudaMemcpyAsync(y, w->Y, sizeof(double) * (size_t)w->B * (N + 1) * n, cudaMemcpyDeviceToHost, w->stream));
  TVB_OK(cudaStreamSynchronize(w->stream));
  return ALTRO_B200_NO_ERROR;
}

// One cached workspace behind the stateless section-B calls, rebuilt when the shape changes
static std::mutex g_cache_mu;
static altro_b200_tvlqr_ws* g_cache = nullptr;
static altro_b200_tvlqr_ws* cached_ws(int batch, int n, int m, int N, bool is_diag) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (g_cache && (g_cache->B != batch || g_cache->n != n || g_cache->m != m || g_cache->N != N ||
                  g_cache->is_diag != (is_diag ? 1 : 0) || g_cache->device != dev)) {
    altro_b200_tvlqr_ws_destroy(g_cache);
    g_cache = nullptr;
  }
  if (!g_cache) g_cache = altro_b200_tvlqr_ws_create(batch, n, m, N, is_diag, dev);
  return g_cache;
}
// -1: shape not compiled in (caller falls back to the run-time-dimension kernel)
int altro_b200_tvlqr_cached_backward(int batch, int n, int m, int N, const double* A, const double* B,
                                     const double* f, const double* Q, const double* R, const double* H,
                                     const double* q, const double* r, double reg, bool is_diag, double* K,
                                     double* d, double* P, double* p, double* delta_V, int* status) {
  if (!find_shape(n, m)) return -1;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  altro_b200_tvlqr_ws* w = cached_ws(batch, n, m, N, is_diag);
  if (!w) return ALTRO_B200_ERR_NO_DEVICE;
  int e = altro_b200_tvlqr_ws_upload(w, A, B, f, Q, R, H, q, r);
  if (!e) e = altro_b200_tvlqr_ws_backward(w, reg);
  if (!e) e = altro_b200_tvlqr_ws_download(w, K, d, P, p, delta_V, status);
  return e;
}
int altro_b200_tvlqr_cached_forward(int batch, int n, int m, int N, const double* A, const double* B,
                                    const double* f, const double* K, const double* d, const double* P,
                                    const double* p, const double* x0, double* x, double* u, double* y) {
  if (!find_shape(n, m)) return -1;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  altro_b200_tvlqr_ws* w = cached_ws(batch, n, m, N, true);
  if (!w) return ALTRO_B200_ERR_NO_DEVICE;
  int e = altro_b200_tvlqr_ws_upload(w, A, B, f, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (!e) e = altro_b200_tvlqr_ws_set_gains(w, K, d, y ? P : nullptr, y ? p : nullptr);
  if (!e) e = altro_b200_tvlqr_ws_forward(w, x0, x, u, y);
  return e;
}

long altro_b200_tvlqr_ws_bytes_per_knot(const altro_b200_tvlqr_ws* w) {
  // compulsory HBM bytes of one backward step of one problem: inputs read + K, d, P, p written
  return w ? 8L * w->rows.kRows : 0;
}

}  // extern "C"
