// solver_team.cuh -- the Riccati sweep with the blocks of ONE problem spread over the warps of a CTA.
//
// k_phase_backward keeps a whole trajectory in one thread: 1 busy warp per group of 32 problems,
// ~650 dependent-ish DFMA per knot for (5,2), and for n = 12 blocks that no longer fit the
// register file (local memory, 5-7 % of the HBM roofline in round 1).  Here a CTA owns a group and
// its W warps split every n-column block BY COLUMNS: warp w owns columns j = w, w + W, ... of
// T1 = A'P+, Qxx, Qux, K, Quu K, P (lane = problem, as everywhere else).  A product
// C[:, j] = op(A) * B[:, j] needs the whole left operand and only the OWN columns of the right one,
// so the warps exchange exactly the blocks that appear on the left -- T1, T2 = B'P+, K, Quu K, Qux
// and the vector p+ -- through shared memory ([element][lane]: conflict-free), two CTA barriers
// per knot.  The small m x m work (Quu, its Cholesky factor, d) is done redundantly by every warp.
// Per warp the knot shrinks from ~650 to ~150 DFMA for (5,2) and the per-thread state to a few
// columns, which is what lets n = 12 stay on chip: P+ columns in registers, [A B] read in place
// from the TMA stage.
//
// Every output element is still produced by ONE thread with the same fma chain, in the same
// order, as TrajSolver::riccati_step, so the results are bit-identical to the single-thread sweep
// and to the persistent twin (tests: pipeline == twin).
#pragma once
#include "solver_kernels.cuh"

namespace altro_b200 {

template <int N_, int M_>
struct TeamShape {
  static constexpr int W = N_ <= 6 ? N_ : 6;          // warps per group
  static constexpr int NC = (N_ + W - 1) / W;         // columns of an n-column block per warp
  // exchange rows: T1 n*n | T2 m*n | K m*n | Quu K m*n | Qux m*n | p n
  static constexpr int kXchRows = N_ * N_ + 4 * M_ * N_ + N_;
};

// element (r, j) of a register-resident R x C block for a column index j that is only known at run
// time (the warp's own column): a chain of selects over static indices keeps the block in
// registers, where M[r + R * j] would force it into local memory
template <int R, int C>
__device__ __forceinline__ double col_pick(const double* M, int r, int j) {
  double v = 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) v = (c == j) ? M[r + R * c] : v;
  return v;
}

// diagonal of the cost Hessian + Gauss-Newton AL terms when every contribution is diagonal
// (diagonal cost, selector rows with linear cones: CON <= 1); same additions in the same order as
// cost_hessian + al_hessian, so the diagonal entries carry the same bits
template <class TS>
__device__ __forceinline__ void hessian_diag(const TS& s, const DeviceProblem& P, int k, bool terminal,
                                             double* dxx, double* duu) {
  constexpr int n = TS::n, m = TS::m;
#pragma unroll
  for (int i = 0; i < n; ++i) dxx[i] = s.Qd[k * n + i];
  if (!terminal) {
#pragma unroll
    for (int i = 0; i < m; ++i) duu[i] = s.Rd[k * m + i];
  }
  if constexpr (TS::kCon != 0) {
    const ConTable& T = P.contab;
    for (int j = 0; j < T.ncon; ++j) {
      const ConSlot& c = T.slot[j];
      if (k < c.k_start || k >= c.k_stop) continue;
      const long zrow = s.zoff(k, c.row0);
      for (int i = 0; i < c.dim; ++i) {
        const int id = c.idx[i];
        if (id < 0) continue;
        const double zt = s.zest_read(zrow, c.row0, i);
        double act = 0.0;
        if (c.cone == CONE_EQUALITY) act = 1.0;
        if (c.cone == CONE_INEQUALITY) act = (zt <= 0) ? 1.0 : 0.0;
        const double gg = s.rho * ((act * c.scale[i]) * (act * c.scale[i]));
#pragma unroll
        for (int e = 0; e < n; ++e) dxx[e] += (id == e) ? gg : 0.0;
        if (!terminal) {
#pragma unroll
          for (int e = 0; e < m; ++e) duu[e] += (id == n + e) ? gg : 0.0;
        }
      }
    }
  }
}

// K1 (team form): CalcExpansions + tvlqr_BackwardPass (solver.cpp:448-449, tvlqr.cpp:65-195) by
// the W warps of the group's CTA, then -- warp 0 alone -- the alpha = 0 half of ForwardPass and
// the start of the line search, exactly as k_phase_backward.
template <class Model, int CON>
__global__ void __launch_bounds__(32 * TeamShape<Model::n, Model::m>::W)
    k_phase_backward_team(const __grid_constant__ DeviceProblem P, int depth, int ring_depth, int wcount,
                          int first) {
  using TS = TrajSolver<Model, CON>;
  using Shape = TeamShape<Model::n, Model::m>;
  constexpr int n = Model::n, m = Model::m, W = Shape::W, NC = Shape::NC;
  constexpr int kV = TS::kV;
  // small blocks: [A B] unpacked into registers; large: read in place from the stage (dense rows)
  constexpr bool kRegJ = (n <= kUnrollDim);
  static_assert(kRegJ || !TS::JP::packed, "in-place Jacobian reads need the dense row layout");
  // full Hessian blocks per thread only where they are cheap or unavoidable (general constraints)
  constexpr bool kFullH = (CON == 2) || kRegJ;
  const int g = P.g0 + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = g * 32 + lane;
  const bool active = b < P.B && (P.flags[b] & TF_ACTIVE);
  if (!__syncthreads_or(active)) return;  // every problem of the group has stopped
  TS s(P, active ? b : g * 32);
  s.rho = (CON && active) ? P.rho[b] : 1.0;

  const int zr = CON ? 2 * P.zrows : 0;
  constexpr int kRowsBw = TS::kRowsBw;  // [J] [lx lu]
  const int sweep_rows = kRowsBw + zr;
  const int scan_rows = TS::kRowsBackwardKernel + zr;
  // shared memory: [pipe barriers 256 B][stages: max(sweep, scan ring)][exchange rows][weights]
  BulkPipe pipe;
  pipe.setup(altro_smem, reinterpret_cast<double*>(altro_smem + 256), depth, sweep_rows * 32);
  const size_t stage_bytes =
      TS::kStaged ? max(BulkPipe::bytes(depth, sweep_rows * 32), 256 + BulkRing::bytes(ring_depth, scan_rows * 32))
                  : BulkPipe::bytes(depth, sweep_rows * 32);
  double* xch = reinterpret_cast<double*>(altro_smem + stage_bytes);
  double* xT1 = xch + lane;
  double* xT2 = xT1 + n * n * 32;
  double* xK = xT2 + m * n * 32;
  double* xQK = xK + m * n * 32;
  double* xQux = xQK + m * n * 32;
  double* xpv = xQux + m * n * 32;
  stage_weights(s, P, xch + Shape::kXchRows * 32, wcount);
  if (tid == 0) pipe.init(W);
  __syncthreads();

  const double* rec = P.xbar + (long)g * P.GS;  // row 0 of the group's knot-0 record
  const double* zrec = CON ? P.z + (long)g * P.GSz : nullptr;
  // pass-local counter c = N - 1 - k (the sweep runs backwards in k)
  auto fetch = [&](int c) {
    const int k = P.N - 1 - c;
    const int st = pipe.acquire(c, (unsigned)sweep_rows * 256u);
    pipe.copy(st, 0, rec + (long)k * P.R + TS::rA * 32, kV * 256);
    pipe.copy(st, kV, rec + (long)k * P.R + TS::rLx * 32, (n + m) * 256);
    if (zr) pipe.copy(st, kRowsBw, zrec + (long)k * P.Rz, zr * 256);
  };
  if (tid == 0)
    for (int c = 0; c < depth && c < P.N; ++c) fetch(c);

  // own columns of the cost-to-go P+ (tvlqr.cpp:85-90 for the terminal knot)
  double Pc[NC][n];
  if (active) {
    double pN[n];
    load_block<n>(s.F(P.lx), s.S, P.N, pN);
    if constexpr (kFullH) {
      double Hxx[n * n];
      s.cost_hessian(P.N, true, Hxx, nullptr, nullptr);
      s.al_hessian(P.N, true, Hxx, nullptr, nullptr);
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = w + jj * W;
        if (j < n) {
#pragma unroll
          for (int i = 0; i < n; ++i) Pc[jj][i] = col_pick<n, n>(Hxx, i, j);
        }
      }
    } else {
      double dxx[n];
      hessian_diag(s, P, P.N, true, dxx, nullptr);
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = w + jj * W;
        if (j < n) {
#pragma unroll
          for (int i = 0; i < n; ++i) Pc[jj][i] = col_pick<1, n>(dxx, 0, i == j ? j : -1);
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < NC; ++jj) {
      const int j = w + jj * W;
      if (j < n) {
        double* Pg = s.F(P.P) + (long)P.N * s.S + (long)n * j * 32;
#pragma unroll
        for (int i = 0; i < n; ++i) Pg[i * 32] = Pc[jj][i];
        const double pNj = col_pick<1, n>(pN, 0, j);
        s.F(P.p)[(long)P.N * s.S + j * 32] = pNj;
        xpv[j * 32] = pNj;
      }
    }
  }

  bool alive = active;
  for (int c = 0; c < P.N; ++c) {
    const int k = P.N - 1 - c;
    const double* st = pipe.wait(c);
    const double* sl = st + lane;  // this lane's column of the stage
    // ---- (b) T1 = A' P+, T2 = B' P+ : own columns                      tvlqr.cpp:135, :139
    double A[kRegJ ? n * n : 1], Bm[kRegJ ? n * m : 1];
    if constexpr (kRegJ) {
      if (alive) s.unstage_jac(st, 0, lane, k, A, Bm);
    }
    // element (l, i) of A / B: registers for the small blocks, in place from the stage otherwise
    // (macros, not lambdas: a by-reference capture would pin the arrays in local memory)
#define ALTRO_AAT(l, i) (kRegJ ? A[kRegJ ? (l) + n * (i) : 0] : sl[((l) + n * (i)) * 32])
#define ALTRO_BAT(l, i) (kRegJ ? Bm[kRegJ ? (l) + n * (i) : 0] : sl[(n * n + (l) + n * (i)) * 32])
    // (one shared-memory operand feeds the NC own columns: the jj loop is innermost everywhere)
    if (alive) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
        double acc[NC];
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) acc[jj] = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) {
          const double a = ALTRO_AAT(l, i);
#pragma unroll
          for (int jj = 0; jj < NC; ++jj) acc[jj] = fma(a, Pc[jj][l], acc[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < NC; ++jj)
          if (w + jj * W < n) xT1[(i + n * (w + jj * W)) * 32] = acc[jj];
      }
#pragma unroll
      for (int i = 0; i < m; ++i) {
        double acc[NC];
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) acc[jj] = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) {
          const double a = ALTRO_BAT(l, i);
#pragma unroll
          for (int jj = 0; jj < NC; ++jj) acc[jj] = fma(a, Pc[jj][l], acc[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < NC; ++jj)
          if (w + jj * W < n) xT2[(i + m * (w + jj * W)) * 32] = acc[jj];
      }
    }
    __syncthreads();  // T1, T2 and p+ complete
    // ---- (c) action-value blocks, gains                                tvlqr.cpp:125-166
    double Quu[m * m], Qu[m], dd[m], pfull[n];
    double Qxxc[NC][n], Quxc[NC][m], Kc[NC][m], QKc[NC][m], Qxj[NC];
    bool ok = true;
    if (alive) {
      if (zr) s.zstage = sl + kRowsBw * 32;
      double Hxx[kFullH ? n * n : 1], Hux[kFullH ? m * n : 1], dxx[kFullH ? 1 : n];
      if constexpr (kFullH) {
        s.cost_hessian(k, false, Hxx, Quu, Hux);
        s.al_hessian(k, false, Hxx, Quu, Hux);
      } else {
        double duu[m];
        hessian_diag(s, P, k, false, dxx, duu);
#pragma unroll
        for (int i = 0; i < m * m; ++i) Quu[i] = 0.0;
#pragma unroll
        for (int i = 0; i < m; ++i) Quu[i + m * i] = duu[i];
      }
      s.zstage = nullptr;
#pragma unroll
      for (int i = 0; i < n; ++i) pfull[i] = xpv[i * 32];
      // Quu += T2 B                                                     :140
#pragma unroll
      for (int j = 0; j < m; ++j)
#pragma unroll
        for (int i = 0; i < m; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < n; ++l) acc = fma(xT2[(i + m * l) * 32], ALTRO_BAT(l, j), acc);
          Quu[i + m * j] += acc;
        }
      // Qu = r + B' p+                                                  :151-152 (f = 0)
#pragma unroll
      for (int i = 0; i < m; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) acc = fma(ALTRO_BAT(l, i), pfull[l], acc);
        Qu[i] = sl[(kV + n + i) * 32] + acc;
      }
      // own columns of A in registers: the products below then take ONE shared-memory operand
      // (T1 / T2 element) per NC multiply-adds
      double Acol[NC][n];
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = (w + jj * W < n) ? w + jj * W : 0;
#pragma unroll
        for (int l = 0; l < n; ++l) {
          if constexpr (kRegJ) Acol[jj][l] = col_pick<n, n>(A, l, j);
          else Acol[jj][l] = sl[(l + n * j) * 32];
        }
      }
      // Qxx[:, j] = Q[:, j] + T1 A[:, j]                                :136
#pragma unroll
      for (int i = 0; i < n; ++i) {
        double acc[NC];
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) acc[jj] = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) {
          const double t = xT1[(i + n * l) * 32];
#pragma unroll
          for (int jj = 0; jj < NC; ++jj) acc[jj] = fma(t, Acol[jj][l], acc[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
          const int j = w + jj * W;
          double h0;
          if constexpr (kFullH) h0 = col_pick<n, n>(Hxx, i, j);
          else h0 = col_pick<1, n>(dxx, 0, i == j ? j : -1);
          Qxxc[jj][i] = h0 + acc[jj];
        }
      }
      // Qux[:, j] = H[:, j] + T2 A[:, j]                                :143
#pragma unroll
      for (int i = 0; i < m; ++i) {
        double acc[NC];
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) acc[jj] = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) {
          const double t = xT2[(i + m * l) * 32];
#pragma unroll
          for (int jj = 0; jj < NC; ++jj) acc[jj] = fma(t, Acol[jj][l], acc[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
          double h0 = 0.0;
          if constexpr (kFullH) h0 = col_pick<m, n>(Hux, i, w + jj * W);
          Quxc[jj][i] = h0 + acc[jj];
        }
      }
      // Qx[j] = q[j] + A[:, j]' p+                                      :147-150
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = w + jj * W;
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) acc = fma(Acol[jj][l], pfull[l], acc);
        Qxj[jj] = sl[(kV + (j < n ? j : 0)) * 32] + acc;
      }
      // K = Qux, d = -Qu, L = chol(Quu)                                  :157-164
      double L[m * m];
#pragma unroll
      for (int i = 0; i < m * m; ++i) L[i] = Quu[i];
#pragma unroll
      for (int i = 0; i < m; ++i) dd[i] = -Qu[i];
      ok = cholesky<m>(L);
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = w + jj * W;
        if (j < n) {
#pragma unroll
          for (int i = 0; i < m; ++i) Kc[jj][i] = Quxc[jj][i];
          if (ok) cholesky_solve<m, 1>(L, Kc[jj]);                     // :165
          double* Kg = s.F(P.K) + (long)k * s.S + (long)m * j * 32;
#pragma unroll
          for (int i = 0; i < m; ++i) Kg[i * 32] = Kc[jj][i];
        }
      }
      if (ok) cholesky_solve<m, 1>(L, dd);                             // :166
      if (w == 0) store_block<m>(s.F(P.d), s.S, k, dd);
      if (ok) {
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
          const int j = w + jj * W;
          if (j < n) {
            // (Quu K)[:, j]                                              :173
#pragma unroll
            for (int i = 0; i < m; ++i) {
              double acc = 0.0;
#pragma unroll
              for (int l = 0; l < m; ++l) acc = fma(Quu[i + m * l], Kc[jj][l], acc);
              QKc[jj][i] = acc;
            }
#pragma unroll
            for (int i = 0; i < m; ++i) {
              xK[(i + m * j) * 32] = Kc[jj][i];
              xQK[(i + m * j) * 32] = QKc[jj][i];
              xQux[(i + m * j) * 32] = Quxc[jj][i];
            }
          }
        }
      }
    }
    // the stage is consumed (in-place [A B], lx, lu and the dual rows were read above)
    pipe.release(c, lane);
    if (w == 0 && c >= 1 && c - 1 + depth < P.N) {
      pipe.wait_writable(c - 1 + depth);
      if (lane == 0) fetch(c - 1 + depth);
    }
    // a failed Cholesky ends the sweep for this problem with the unsolved K = Qux, d = -Qu stored
    // and everything below left stale (tvlqr.cpp:162-164, ignored by Solve: quirk Q2)
    if (!ok) alive = false;
    __syncthreads();  // K, Quu K, Qux complete
    // ---- (d) cost-to-go: own columns                                   tvlqr.cpp:173-186
    if (alive) {
      double Pj[NC][n];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        double a1[NC], a2[NC], a3[NC];
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) a1[jj] = a2[jj] = a3[jj] = 0.0;
#pragma unroll
        for (int l = 0; l < m; ++l) {
          const double qk = xQK[(l + m * i) * 32], kk = xK[(l + m * i) * 32], qq = xQux[(l + m * i) * 32];
#pragma unroll
          for (int jj = 0; jj < NC; ++jj) {
            a1[jj] = fma(qk, Kc[jj][l], a1[jj]);    // ((Quu K)' K)[i, j]
            a2[jj] = fma(kk, Quxc[jj][l], a2[jj]);  // (K' Qux)[i, j]
            a3[jj] = fma(Kc[jj][l], qq, a3[jj]);    // (K' Qux)[j, i]
          }
        }
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
          double v = Qxxc[jj][i] + a1[jj];
          v -= a2[jj];
          v -= a3[jj];
          Pj[jj][i] = v;
        }
      }
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = w + jj * W;
        if (j < n) {
          double pj = Qxj[jj];
          {
            double b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
            for (int l = 0; l < m; ++l) b1 = fma(QKc[jj][l], dd[l], b1);
#pragma unroll
            for (int l = 0; l < m; ++l) b2 = fma(Kc[jj][l], Qu[l], b2);
#pragma unroll
            for (int l = 0; l < m; ++l) b3 = fma(Quxc[jj][l], dd[l], b3);
            pj -= b1;
            pj -= b2;
            pj += b3;
          }
          double* Pg = s.F(P.P) + (long)k * s.S + (long)n * j * 32;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            Pg[i * 32] = Pj[jj][i];
            Pc[jj][i] = Pj[jj][i];
          }
          s.F(P.p)[(long)k * s.S + j * 32] = pj;
          xpv[j * 32] = pj;
        }
      }
    }
  }
  // K, d were written through the generic proxy by all warps; warp 0's scans read them with bulk
  // copies (async proxy)
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  if (w != 0) return;
  double phi0 = 0.0, dphi0 = 0.0;
  if constexpr (TS::kStaged) {
    BulkRing ring;
    ring.init(altro_smem + 256, ring_depth, scan_rows * 32, lane == 0);
    __syncwarp();
    backward_scans<Model, CON>(P, s, ring, ring_depth, g, lane, b, active, first, phi0, dphi0);
  } else {
    // large blocks: a scan stage would not fit next to the exchange rows; direct loads
    if (active) s.phase_phi0_scan(&phi0, &dphi0);
  }
  backward_finish(P, b, active, phi0, dphi0);
}

#undef ALTRO_AAT
#undef ALTRO_BAT

}  // namespace altro_b200
