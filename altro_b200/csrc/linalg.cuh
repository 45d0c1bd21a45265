// linalg.cuh -- fixed-size dense helpers for the thread-per-trajectory kernels.
//
// Every block the Riccati recursion touches is tiny (n <= 12, m <= 4), so each
// trajectory keeps its blocks in registers of ONE thread and a warp works on the 32 problems of
// one group of the knot-record HBM layout (one coalesced 256-byte row per matrix element).
// All matrices are column-major (Eigen's default, which the reference's raw-pointer API exposes:
// altro_solver.hpp:185) and sizes are template parameters so loops unroll into straight-line
// DFMA code with no indexing.
#pragma once

namespace altro_b200 {

// blocks up to this dimension are fully unrolled; larger ones keep rolled loops so the
// instruction footprint stays inside the I-cache and arrays live in local memory
constexpr int kUnrollDim = 6;

#define ALTRO_DEV __device__ __forceinline__

// C (RA x CB) (=, +=, -=) op(A) * op(B);  op(A) is RA x KK, op(B) is KK x CB.
// ACC: 0 assign, 1 add, -1 subtract.
template <int RA, int CB, int KK, bool TA, bool TB, int ACC>
ALTRO_DEV void mm(const double* __restrict__ A, const double* __restrict__ B, double* C) {
  constexpr int lda = TA ? KK : RA;
  constexpr int ldb = TB ? CB : KK;
  constexpr bool full = (RA <= kUnrollDim && CB <= kUnrollDim && KK <= kUnrollDim);
  if constexpr (full) {
#pragma unroll
    for (int j = 0; j < CB; ++j) {
#pragma unroll
      for (int i = 0; i < RA; ++i) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < KK; ++l) {
          const double a = TA ? A[l + lda * i] : A[i + lda * l];
          const double b = TB ? B[j + ldb * l] : B[l + ldb * j];
          s = fma(a, b, s);
        }
        if (ACC == 0) C[i + RA * j] = s;
        if (ACC == 1) C[i + RA * j] += s;
        if (ACC == -1) C[i + RA * j] -= s;
      }
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < CB; ++j) {
#pragma unroll 1
      for (int i = 0; i < RA; ++i) {
        double s = 0.0;
#pragma unroll 4
        for (int l = 0; l < KK; ++l) {
          const double a = TA ? A[l + lda * i] : A[i + lda * l];
          const double b = TB ? B[j + ldb * l] : B[l + ldb * j];
          s = fma(a, b, s);
        }
        if (ACC == 0) C[i + RA * j] = s;
        if (ACC == 1) C[i + RA * j] += s;
        if (ACC == -1) C[i + RA * j] -= s;
      }
    }
  }
}

template <int L>
ALTRO_DEV double dot(const double* a, const double* b) {
  double s = 0.0;
  if constexpr (L <= 2 * kUnrollDim) {
#pragma unroll
    for (int i = 0; i < L; ++i) s = fma(a[i], b[i], s);
  } else {
#pragma unroll 4
    for (int i = 0; i < L; ++i) s = fma(a[i], b[i], s);
  }
  return s;
}

template <int L>
ALTRO_DEV double infnorm(const double* a) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < L; ++i) s = fmax(s, fabs(a[i]));
  return s;
}

// In-place lower Cholesky of an M x M block.  Fails (returns false) on a non-positive pivot,
// which is when Eigen's LLT used by the reference reports NumericalIssue (tvlqr.cpp:161-164).
template <int M>
ALTRO_DEV bool cholesky(double* Q) {
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double x = Q[j + M * j];
#pragma unroll
    for (int l = 0; l < j; ++l) x = fma(-Q[j + M * l], Q[j + M * l], x);
    if (x <= 0.0) return false;
    x = sqrt(x);
    Q[j + M * j] = x;
    const double inv = 1.0 / x;
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      double s = Q[i + M * j];
#pragma unroll
      for (int l = 0; l < j; ++l) s = fma(-Q[i + M * l], Q[j + M * l], s);
      Q[i + M * j] = s * inv;
    }
  }
  return true;
}

// X (M x C) <- (L L^T)^-1 X
template <int M, int C>
ALTRO_DEV void cholesky_solve(const double* L, double* X) {
  double inv[M];
#pragma unroll
  for (int i = 0; i < M; ++i) inv[i] = 1.0 / L[i + M * i];
  constexpr bool full = (C <= kUnrollDim);
  auto solve_col = [&](double* x) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double s = x[i];
#pragma unroll
      for (int l = 0; l < i; ++l) s = fma(-L[i + M * l], x[l], s);
      x[i] = s * inv[i];
    }
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
      double s = x[i];
#pragma unroll
      for (int l = i + 1; l < M; ++l) s = fma(-L[l + M * i], x[l], s);
      x[i] = s * inv[i];
    }
  };
  if constexpr (full) {
#pragma unroll
    for (int c = 0; c < C; ++c) solve_col(X + M * c);
  } else {
#pragma unroll 1
    for (int c = 0; c < C; ++c) solve_col(X + M * c);
  }
}

// Knot-record field access (device_problem.h): `base` already points at this problem's lane of
// the field's first row in knot 0 of its group; knot k is `rec` doubles further, rows of a block
// are 32 doubles (one 256-byte line holding the 32 problems of the group) apart.
constexpr int kLanes = 32;
template <int E>
ALTRO_DEV void load_block(const double* __restrict__ base, long rec, int k, double* out) {
  const double* p = base + (long)k * rec;
  if constexpr (E <= 4 * kUnrollDim * kUnrollDim) {
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = p[e * kLanes];
  } else {
#pragma unroll 8
    for (int e = 0; e < E; ++e) out[e] = p[e * kLanes];
  }
}

// Software prefetch of the E rows of knot k into L2 (no register cost): used by the sweeps that
// do not stage through shared memory.
template <int E>
ALTRO_DEV void prefetch_block(const double* __restrict__ base, long rec, int k) {
  const double* p = base + (long)k * rec;
  if constexpr (E <= 4 * kUnrollDim * kUnrollDim) {
#pragma unroll
    for (int e = 0; e < E; ++e) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + e * kLanes));
  } else {
#pragma unroll 8
    for (int e = 0; e < E; ++e) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + e * kLanes));
  }
}

// ---- asynchronous staging of the NEXT knots' rows into shared memory (cp.async / LDGSTS).
// The sequential sweeps are bound by the latency of their per-knot loads (ncu r01 v2: 60-70 % of
// all stall cycles are long-scoreboard even with an L2 prefetch one knot ahead).  Each lane copies
// the elements of ITS OWN problem into ITS OWN column of a per-warp ring in shared memory,
//     ring[stage][element][lane]      (256-byte rows: conflict-free 8-byte accesses)
// one to three knots ahead, and reads them back with LDS (~30 cycles) when the knot's turn comes.
// A lane only ever reads what it copied itself, so cp.async.wait_group is the only
// synchronisation needed -- no barrier -- and compacted (non-contiguous) problem lists work.
ALTRO_DEV void cp_async_f64(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
ALTRO_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` (0..2) of the most recent groups are still in flight
ALTRO_DEV void cp_async_wait(int pending) {
  if (pending <= 0)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  else if (pending == 1)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
  else
    asm volatile("cp.async.wait_group 2;" ::: "memory");
}
constexpr int kMaxStageDepth = 3;

// stage: st points at this lane's column of one ring stage; `off` = first element row
template <int E>
ALTRO_DEV void stage_block(double* st, int off, const double* __restrict__ base, long rec, int k) {
  const double* p = base + (long)k * rec;
#pragma unroll
  for (int e = 0; e < E; ++e) cp_async_f64(st + (off + e) * 32, p + e * 32);
}
template <int E>
ALTRO_DEV void unstage_block(const double* st, int off, double* out) {
#pragma unroll
  for (int e = 0; e < E; ++e) out[e] = st[(off + e) * 32];
}

template <int E>
ALTRO_DEV void store_block(double* __restrict__ base, long rec, int k, const double* in) {
  double* p = base + (long)k * rec;
  if constexpr (E <= 4 * kUnrollDim * kUnrollDim) {
#pragma unroll
    for (int e = 0; e < E; ++e) p[e * kLanes] = in[e];
  } else {
#pragma unroll 8
    for (int e = 0; e < E; ++e) p[e * kLanes] = in[e];
  }
}

}  // namespace altro_b200
