// linalg.cuh -- fixed-size dense helpers for the thread-per-trajectory kernels.
//
// Every block the Riccati recursion touches is tiny (n <= 12, m <= 4), so each
// trajectory keeps its blocks in registers of ONE thread and a warp works on the 32 problems of
// one group of the knot-record HBM layout (one coalesced 256-byte row per matrix element).
// All matrices are column-major (Eigen's default, which the reference's raw-pointer API exposes:
// altro_solver.hpp:185) and sizes are template parameters so loops unroll into straight-line
// DFMA code with no indexing.
#pragma once
#include <cstdio>

namespace altro_b200 {

// blocks up to this dimension are fully unrolled; larger ones keep rolled loops so the
// instruction footprint stays inside the I-cache and arrays live in local memory
constexpr int kUnrollDim = 6;

#define ALTRO_DEV __device__ __forceinline__

// register-tile edge for the large-block path of mm: the largest of 6, 4, 3, 2, 1 dividing d
template <int D>
struct TileDim {
  static constexpr int v = D % 6 == 0 ? 6 : (D % 4 == 0 ? 4 : (D % 3 == 0 ? 3 : (D % 2 == 0 ? 2 : 1)));
};

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
    // Large blocks: operands live in local memory, so a rolled dot product costs two loads per
    // FMA.  Register tiles of TR x TC accumulators cut that to (TR + TC) loads per TR * TC FMAs;
    // every output element still accumulates over l in ascending order, i.e. the result is
    // bit-identical to the plain triple loop.
    constexpr int TR = TileDim<RA>::v, TC = TileDim<CB>::v;
#pragma unroll 1
    for (int j0 = 0; j0 < CB; j0 += TC) {
#pragma unroll 1
      for (int i0 = 0; i0 < RA; i0 += TR) {
        double acc[TR * TC];
#pragma unroll
        for (int t = 0; t < TR * TC; ++t) acc[t] = 0.0;
#pragma unroll 2
        for (int l = 0; l < KK; ++l) {
          double av[TR], bv[TC];
#pragma unroll
          for (int i = 0; i < TR; ++i) av[i] = TA ? A[l + lda * (i0 + i)] : A[(i0 + i) + lda * l];
#pragma unroll
          for (int j = 0; j < TC; ++j) bv[j] = TB ? B[(j0 + j) + ldb * l] : B[l + ldb * (j0 + j)];
#pragma unroll
          for (int j = 0; j < TC; ++j)
#pragma unroll
            for (int i = 0; i < TR; ++i) acc[i + TR * j] = fma(av[i], bv[j], acc[i + TR * j]);
        }
#pragma unroll
        for (int j = 0; j < TC; ++j)
#pragma unroll
          for (int i = 0; i < TR; ++i) {
            const double sum = acc[i + TR * j];
            if (ACC == 0) C[(i0 + i) + RA * (j0 + j)] = sum;
            if (ACC == 1) C[(i0 + i) + RA * (j0 + j)] += sum;
            if (ACC == -1) C[(i0 + i) + RA * (j0 + j)] -= sum;
          }
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

// ---- TMA staging of whole knot records into shared memory (cp.async.bulk, SASS UBLKCP).
// The sequential sweeps read, per knot, one contiguous multi-KB range of the group's knot record
// (device_problem.h).  A BulkRing keeps the next `depth` knots in flight: the leader thread arms
// the stage's mbarrier with the byte count and issues one bulk copy per contiguous piece; every
// thread waits on the mbarrier parity, reads ITS lane's column of the stage
//     stage[row][lane]      (256-byte rows: conflict-free 8-byte accesses)
// into registers, WORKS ON THE KNOT, and only then passes the warp/CTA barrier after which the
// leader refills the stage: a barrier does not wait for outstanding LDS, but arithmetic that
// consumed the values does, so the bulk copy can never overwrite a row that is still being read
// (with 10 warps x 27 LDS queued per knot the read-out takes as long as an L2-hit bulk copy; run
// to run differences in the last bits showed up before the order was fixed).  The refill
// therefore runs depth - 1 knots ahead.  Measured at the
// 3.5 warps/SM of B = 16384: 6.8 TB/s vs 3.2 TB/s for per-element LDG from the old
// problem-fastest layout and 5.0 TB/s for per-lane cp.async (profiles/r01_microbench_layout.txt).
constexpr int kMaxStageDepth = 4;

struct BulkRing {
  double* data;             // [depth][stage_doubles], 128-byte aligned
  unsigned long long* bar;  // [depth] "stage full" mbarriers
  int depth, stage_doubles;
  int s;                    // stage the consumer reads next
  unsigned parity;

  // smem: dynamic shared memory of the CTA (>= 128 + depth * stage_doubles * 8 bytes).  All threads
  // call init; the caller synchronises the CTA (or warp) afterwards.
  ALTRO_DEV void init(unsigned char* smem, int depth_, int stage_doubles_, bool leader) {
    bar = reinterpret_cast<unsigned long long*>(smem);
    data = reinterpret_cast<double*>(smem + 128);
    depth = depth_;
    stage_doubles = stage_doubles_;
    s = 0;
    parity = 0;
    if (leader) {
      for (int j = 0; j < depth; ++j) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar + j);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // leader only: announce `bytes` for `stage`, then issue the copies that add up to it
  ALTRO_DEV void expect(int stage, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar + stage);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
  }
  ALTRO_DEV void copy(int stage, int row, const double* src, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar + stage);
    const unsigned d = (unsigned)__cvta_generic_to_shared(data + (long)stage * stage_doubles + row * 32);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(src), "r"(bytes), "r"(a)
        : "memory");
  }
  // all threads: block until the current stage has landed; returns its first row
  ALTRO_DEV const double* wait() const {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar + s);
    unsigned ok = 0;
    unsigned long long spins = 0;
    while (!ok) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(ok)
          : "r"(a), "r"(parity)
          : "memory");
      if (!ok && ++spins > (1ull << 21)) {  // a protocol bug must fail the launch, not hang the device
        printf("altro_b200: ring wait stuck (block %d thread %d stage %d parity %u)\n", (int)blockIdx.x,
               (int)threadIdx.x, s, parity);
        __trap();
      }
    }
    return data + (long)s * stage_doubles;
  }
  ALTRO_DEV void advance() {
    if (++s == depth) {
      s = 0;
      parity ^= 1u;
    }
  }
  __host__ __device__ static size_t bytes(int depth, int stage_doubles) {
    return 128 + (size_t)depth * stage_doubles * 8;
  }
};

// ---- The same staging for SEVERAL consumer warps that are NOT kept in lockstep (k_phase_forward,
// k_phase_backward_team): a classic full/empty mbarrier pipeline.  `full[s]` completes when the
// bulk copies of a stage have landed (1 arrival + transaction bytes); `empty[s]` completes when
// every consumer warp has released the stage (one elected arrival per warp, issued after the warp
// consumed what it read).  The producer thread refills a stage only after its `empty` phase
// completed, so fast warps run up to depth - 1 knots ahead of slow ones instead of meeting at a
// CTA barrier every knot (ncu r01: 9 barrier stalls per issue in the late line-search rounds).
// A pass = one sweep over the knots.  The barriers are initialised ONCE per kernel with a fixed
// consumer count and never invalidated: stage and parity of pass-local knot k follow from the
// running count c0 of knots the pipe has carried in earlier passes.  (Re-arming the barriers per
// pass with mbarrier.inval + init lost arrivals as soon as two CTAs shared an SM -- found on the
// B200 with the watchdog below; every warp of the CTA is therefore a consumer of every pass, a warp
// with nothing to compute just waits and releases.)
struct BulkPipe {
  unsigned long long* full;   // [depth]
  unsigned long long* empty;  // [depth]
  double* data;               // [depth][stage_doubles], 128-byte aligned
  int depth, stage_doubles;
  int c0;                     // knots carried by earlier passes

  static constexpr int kBarBytes = 128;  // barrier block of one pipe: full[<=8] | empty[<=8]
  __host__ __device__ static size_t bytes(int depth, int stage_doubles) {
    return 256 + (size_t)depth * stage_doubles * 8;
  }
  // pointer carving only; all threads.  bars: kBarBytes of shared memory, data: the stages
  ALTRO_DEV void setup(unsigned char* bars, double* data_, int depth_, int stage_doubles_) {
    full = reinterpret_cast<unsigned long long*>(bars);
    empty = full + 8;
    data = data_;
    depth = depth_;
    stage_doubles = stage_doubles_;
    c0 = 0;
  }
  // ONE thread, once per kernel, followed by a CTA barrier
  ALTRO_DEV void init(int consumer_warps) const {
    for (int j = 0; j < depth; ++j) {
      const unsigned f = (unsigned)__cvta_generic_to_shared(full + j);
      const unsigned e = (unsigned)__cvta_generic_to_shared(empty + j);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(e), "r"(consumer_warps) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // every thread that tracks the pipe, after a pass of `knots` knots
  ALTRO_DEV void end_pass(int knots) { c0 += knots; }

  // A wait that cannot complete is a protocol bug; it must surface as a kernel error (trap -> the
  // launch fails, the C ABI returns an error), never as a hung device.
  ALTRO_DEV static void spin(unsigned addr, unsigned parity, int what = 0, int k = -1) {
    unsigned ok = 0;
    unsigned long long spins = 0;
    while (!ok) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(ok)
          : "r"(addr), "r"(parity)
          : "memory");
      if (!ok && ++spins > (1ull << 22)) {
        if ((threadIdx.x & 31) == 0 || what == 2)
          printf("altro_b200: mbarrier wait stuck: %s knot %d (block %d of %d, thread %d of %d, addr %u parity %u)\n",
                 what == 0 ? "full" : (what == 1 ? "empty/warp" : "empty/producer"), k, (int)blockIdx.x,
                 (int)gridDim.x, (int)threadIdx.x, (int)blockDim.x, addr, parity);
        __trap();
      }
    }
  }
  // ALL lanes of the producer's warp, before the producer thread calls acquire(k): wait until every
  // consumer warp has released the stage's previous use.  Waiting as a whole warp keeps the
  // producer thread converged with its warp (a lone spinning lane would make the warp execute the
  // next knot twice: once without it, once for it).
  ALTRO_DEV void wait_writable(int k) const {
    const int g = c0 + k;
    const int use = g / depth;
    if (use > 0) spin((unsigned)__cvta_generic_to_shared(empty + g % depth), (unsigned)((use - 1) & 1), 1, k);
  }
  // producer thread: make the stage of pass-local knot k writable, announce `bytes`; returns the stage
  ALTRO_DEV int acquire(int k, unsigned bytes) const {
    const int g = c0 + k;
    const int st = g % depth;
    const int use = g / depth;
    if (use > 0) spin((unsigned)__cvta_generic_to_shared(empty + st), (unsigned)((use - 1) & 1), 2, k);
    const unsigned a = (unsigned)__cvta_generic_to_shared(full + st);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
    return st;
  }
  ALTRO_DEV void copy(int stage, int row, const double* src, unsigned bytes) const {
    const unsigned a = (unsigned)__cvta_generic_to_shared(full + stage);
    const unsigned d = (unsigned)__cvta_generic_to_shared(data + (long)stage * stage_doubles + row * 32);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(src), "r"(bytes), "r"(a)
        : "memory");
  }
  // consumer, all lanes: wait for pass-local knot k; returns the stage's first row
  ALTRO_DEV const double* wait(int k) const {
    const int g = c0 + k;
    const int st = g % depth;
    spin((unsigned)__cvta_generic_to_shared(full + st), (unsigned)((g / depth) & 1), 0, k);
    return data + (long)st * stage_doubles;
  }
  // consumer, all lanes of the warp, AFTER the arithmetic that consumed the stage's values
  ALTRO_DEV void release(int k, int lane) const {
    __syncwarp();
    if (lane == 0) {
      const unsigned a = (unsigned)__cvta_generic_to_shared(empty + ((c0 + k) % depth));
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
    }
  }
};

// ---- The staging pipeline of k_phase_forward's rollout passes.  Like BulkPipe, but built for a
// critical-path consumer (the rollout warp), measured with per-warp clocks on the B200 (r02o-r02r:
// the rollout warp spent a third of every knot in pipeline bookkeeping):
//   * no `empty` barrier and no producer that waits: every consumer warp counts itself out of a
//     stage (lane 0, one shared-memory atomic whose result is looked at one knot LATER, so its
//     latency is never exposed) and the warp that counted last refills the stage;
//   * the number of consumers is a per-pass argument: warps without work sit the pass out instead
//     of waiting and releasing every knot (their spin loops took a third of the kernel's issued
//     instructions);
//   * stage index and phase parity are kept incrementally (pass-local stage = k mod depth advanced
//     by the caller, one parity bit per stage) -- no integer division by the run-time depth.
struct RollPipe {
  unsigned long long* full;  // [<= 8] "stage landed" mbarriers: one arrival + the copies' bytes
  unsigned* cnt;             // [<= 8] consumer warps that released the stage's current use
  double* data;              // [depth][stage_doubles], 128-byte aligned
  int depth, stage_doubles;
  unsigned parbits;          // bit s: the parity the next wait on stage s expects
  unsigned passmask;         // stages that are used an odd number of times by a pass of N knots

  // pointer carving only; all threads.  bars: 128 bytes (full[8] | cnt[8] + padding)
  ALTRO_DEV void setup(unsigned char* bars, double* data_, int depth_, int stage_doubles_, int N) {
    full = reinterpret_cast<unsigned long long*>(bars);
    cnt = reinterpret_cast<unsigned*>(full + 8);
    data = data_;
    depth = depth_;
    stage_doubles = stage_doubles_;
    parbits = 0u;
    passmask = 0u;
    for (int s = 0; s < depth && s < N; ++s)
      if (((N - s + depth - 1) / depth) & 1) passmask |= 1u << s;
  }
  ALTRO_DEV void init() const {  // ONE thread, once per kernel, followed by a CTA barrier
    for (int j = 0; j < depth; ++j) {
      const unsigned f = (unsigned)__cvta_generic_to_shared(full + j);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f) : "memory");
      cnt[j] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // whoever refills stage st: announce `bytes`, then issue the copies that add up to it
  ALTRO_DEV void arm(int st, unsigned bytes) const {
    const unsigned a = (unsigned)__cvta_generic_to_shared(full + st);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
  }
  ALTRO_DEV void copy(int st, int row, const double* src, unsigned bytes) const {
    const unsigned a = (unsigned)__cvta_generic_to_shared(full + st);
    const unsigned d = (unsigned)__cvta_generic_to_shared(data + (long)st * stage_doubles + row * 32);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(src), "r"(bytes), "r"(a)
        : "memory");
  }
  // consumer, all lanes: wait for the next use of stage st; returns the stage's first row
  ALTRO_DEV double* wait(int st) {
    mbar_wait((unsigned)__cvta_generic_to_shared(full + st), (parbits >> st) & 1u);
    parbits ^= 1u << st;
    return data + (long)st * stage_doubles;
  }
  // a warp that sits a pass out keeps its parities in step
  ALTRO_DEV void skip_pass() { parbits ^= passmask; }
  // lane 0 of a consumer warp, after the warp consumed the stage (behind a __syncwarp): count out;
  // the returned value says, when it equals consumers - 1, that this warp was the last one
  ALTRO_DEV unsigned count_out(int st) const {
    __threadfence_block();
    return atomicAdd(cnt + st, 1u);
  }
  ALTRO_DEV void reset(int st) const {
    *reinterpret_cast<volatile unsigned*>(cnt + st) = 0u;
    __threadfence_block();
  }
  // A wait that cannot complete is a protocol bug; it must surface as a kernel error (trap), never
  // as a hung device.  try_wait suspends the thread in hardware for up to the hinted time, so the
  // loop around it turns over rarely.
  ALTRO_DEV static void mbar_wait(unsigned addr, unsigned parity) {
    unsigned ok = 0;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    unsigned spins = 0;
    while (!ok) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
          : "=r"(ok)
          : "r"(addr), "r"(parity), "r"(20000u)
          : "memory");
      if (!ok && ++spins > (1u << 20)) {
        if ((threadIdx.x & 31) == 0)
          printf("altro_b200: stage wait stuck (block %d of %d, thread %d of %d, addr %u parity %u)\n",
                 (int)blockIdx.x, (int)gridDim.x, (int)threadIdx.x, (int)blockDim.x, addr, parity);
        __trap();
      }
    }
  }
};

// this lane's element e of a block that starts at `row` of a landed stage
template <int E>
ALTRO_DEV void unstage_block(const double* stage, int row, int lane, double* out) {
#pragma unroll
  for (int e = 0; e < E; ++e) out[e] = stage[(row + e) * 32 + lane];
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
