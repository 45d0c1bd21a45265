"""GPU tests of sections A/B of the C ABI (tvlqr drop-in) and of the C++ facade."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import altro_b200
from test_oracle_cones_tvlqr import (D0_EXPECTED, K0_EXPECTED, XN_EXPECTED, YN_EXPECTED, run_tvlqr,
                                     tvlqr_problem)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


def _cm(M):
    return np.asarray(M, dtype=float).T.reshape(-1).copy()


def gpu_tvlqr_single(pr, is_diag=True, reg=0.0, with_q=False):
    """tvlqr_BackwardPass / tvlqr_ForwardPass through the reference's own pointer-table signature."""
    L = altro_b200.load_library()
    n, m, N = pr["n"], pr["m"], pr["N"]

    def table(arrs):
        keep = [np.ascontiguousarray(a, dtype=float) for a in arrs]
        return (dp * len(keep))(*[a.ctypes.data_as(dp) for a in keep]), keep

    A, kA = table([_cm(pr["A"])] * N)
    Bt, kB = table([_cm(pr["B"])] * N)
    f, kf = table([pr["f"]] * N)
    if is_diag:
        Q, kQ = table([pr["Qd"]] * N + [pr["Qfd"]])
        R, kR = table([pr["Rd"]] * N)
    else:
        Q, kQ = table([_cm(np.diag(pr["Qd"]))] * N + [_cm(np.diag(pr["Qfd"]))])
        R, kR = table([_cm(np.diag(pr["Rd"]))] * N)
    H, kH = table([np.zeros(m * n)] * N)
    q, kq = table([pr["q"]] * (N + 1))
    r, kr = table([pr["r"]] * N)
    outs = {}
    for name, size, cnt in [("K", m * n, N), ("d", m, N), ("P", n * n, N + 1), ("p", n, N + 1),
                            ("x", n, N + 1), ("u", m, N), ("y", n, N + 1), ("Qxx", n * n, N), ("Quu", m * m, N),
                            ("Qux", m * n, N), ("Qx", n, N), ("Qu", m, N)]:
        outs[name] = table([np.zeros(size) for _ in range(cnt)])
    nx = (C.c_int * (N + 1))(*([n] * (N + 1)))
    nu = (C.c_int * N)(*([m] * N))
    dV = np.zeros(2)
    T = lambda k: outs[k][0]
    NULL = None
    L.tvlqr_BackwardPass.restype = C.c_int
    res = L.tvlqr_BackwardPass(nx, nu, N, A, Bt, f, Q, R, H, q, r, C.c_double(reg), T("K"), T("d"),
                               T("P"), T("p"), dV.ctypes.data_as(dp),
                               T("Qxx") if with_q else NULL, T("Quu") if with_q else NULL,
                               T("Qux") if with_q else NULL, T("Qx") if with_q else NULL,
                               T("Qu") if with_q else NULL,
                               NULL, NULL, NULL, NULL, NULL, C.c_bool(False), C.c_bool(is_diag))
    x0 = np.ascontiguousarray(pr["x0"])
    res2 = L.tvlqr_ForwardPass(nx, nu, N, A, Bt, f, T("K"), T("d"), T("P"), T("p"),
                               x0.ctypes.data_as(dp), T("x"), T("u"), T("y"))
    return res, res2, outs, dV


def test_tvlqr_dropin_goldens():
    """tvlqr_test.cpp:185-213 through the GPU drop-in: K0, d0 (1e-6), xN (1e-6), yN (1e-5)."""
    pr = tvlqr_problem(np.float32(0.01))
    res, res2, outs, dV = gpu_tvlqr_single(pr)
    assert res == -1 and res2 == -1
    n, m, N = pr["n"], pr["m"], pr["N"]
    assert np.linalg.norm(outs["K"][1][0].reshape(n, m).T - K0_EXPECTED) < 1e-6
    assert np.linalg.norm(outs["d"][1][0] - D0_EXPECTED) < 1e-6
    assert np.abs(outs["x"][1][N] - XN_EXPECTED).max() < 1e-6
    assert np.abs(outs["y"][1][N] - YN_EXPECTED).max() < 1e-5


def test_tvlqr_dropin_matches_oracle_all_outputs(oracle):
    """Every output of the reference's signature, incl. the action-value expansion tables Qxx, Quu,
    Qux, Qx, Qu (tvlqr.cpp:123-152), with and without the caller passing those tables."""
    for is_diag in (True, False):
        for with_q in (False, True):
            pr = tvlqr_problem(np.float32(0.01))
            res, res2, outs, dV = gpu_tvlqr_single(pr, is_diag, with_q=with_q)
            ores, _, oK0, od0, oxN, oyN, odV, oouts = run_tvlqr(oracle, pr, is_diag)
            assert res == ores
            names = ("K", "d", "P", "p", "x", "u", "y") + (("Qxx", "Quu", "Qux", "Qx", "Qu") if with_q else ())
            for name in names:
                for k in range(len(outs[name][1])):
                    assert np.allclose(outs[name][1][k], oouts[name][1][k], rtol=1e-12, atol=1e-12), (name, k)
            assert np.allclose(dV, odV, rtol=1e-12)


def test_tvlqr_cholesky_failure_convention(oracle):
    """Returns the failing knot index and leaves lower knots untouched (tvlqr.cpp:162-164)."""
    pr = tvlqr_problem(np.float32(0.01))
    pr["Rd"] = np.full(2, -1e6)
    res, _, outs, _ = gpu_tvlqr_single(pr)
    assert res == pr["N"] - 1
    assert np.all(outs["K"][1][0] == 0)      # knot 0 was never reached


def test_tvlqr_batched_random_dims(oracle):
    """Section B on random LQ problems of the sweep dimensions, vs the oracle problem by problem."""
    L = altro_b200.load_library()
    rng = np.random.default_rng(3)
    for n, m, N, B in [(4, 2, 20, 37), (6, 4, 15, 33), (12, 4, 9, 5), (2, 1, 30, 64)]:
        A = np.eye(n) + 0.1 * rng.normal(size=(B, N, n, n))
        Bm = rng.normal(size=(B, N, n, m))
        f = 0.1 * rng.normal(size=(B, N, n))
        def spd(k, cnt):
            M = rng.normal(size=(B, cnt, k, k))
            return M @ np.swapaxes(M, -1, -2) + np.eye(k)
        Q, R = spd(n, N + 1), spd(m, N)
        H = 0.01 * rng.normal(size=(B, N, m, n))
        q, r = rng.normal(size=(B, N + 1, n)), rng.normal(size=(B, N, m))
        cm = lambda M: np.ascontiguousarray(np.swapaxes(M, -1, -2))   # column-major blocks
        K = np.zeros((B, N, n, m)); d = np.zeros((B, N, m)); P = np.zeros((B, N + 1, n, n))
        p = np.zeros((B, N + 1, n)); dV = np.zeros((B, 2)); st = np.zeros(B, dtype=np.int32)
        ptr = lambda a: a.ctypes.data_as(dp)
        Ac, Bc, Qc, Rc, Hc = cm(A), cm(Bm), cm(Q), cm(R), cm(H)
        rc = L.altro_b200_tvlqr_backward_batch(B, n, m, N, ptr(Ac), ptr(Bc), ptr(f), ptr(Qc), ptr(Rc),
                                               ptr(Hc), ptr(q), ptr(r), 0.0, False, ptr(K), ptr(d),
                                               ptr(P), ptr(p), ptr(dV), st.ctypes.data_as(C.POINTER(C.c_int)))
        assert rc == 0 and np.all(st == -1)
        x0 = rng.normal(size=(B, n))
        X = np.zeros((B, N + 1, n)); U = np.zeros((B, N, m)); Y = np.zeros((B, N + 1, n))
        rc = L.altro_b200_tvlqr_forward_batch(B, n, m, N, ptr(Ac), ptr(Bc), ptr(f), ptr(K), ptr(d),
                                              ptr(P), ptr(p), ptr(x0), ptr(X), ptr(U), ptr(Y))
        assert rc == 0
        # oracle, problem by problem, through its pointer-table API
        OL = oracle.lib()
        for b in range(0, B, max(1, B // 6)):
            def tab(arrs):
                keep = [np.ascontiguousarray(a, dtype=float).reshape(-1) for a in arrs]
                return (dp * len(keep))(*[a.ctypes.data_as(dp) for a in keep]), keep
            tA, k1 = tab(list(Ac[b])); tB, k2 = tab(list(Bc[b])); tf, k3 = tab(list(f[b]))
            tQ, k4 = tab(list(Qc[b])); tR, k5 = tab(list(Rc[b])); tH, k6 = tab(list(Hc[b]))
            tq, k7 = tab(list(q[b])); tr, k8 = tab(list(r[b]))
            o = {}
            for name, size, cnt in [("K", m * n, N), ("d", m, N), ("P", n * n, N + 1), ("p", n, N + 1),
                                    ("Qxx", n * n, N), ("Quu", m * m, N), ("Qux", m * n, N), ("Qx", n, N),
                                    ("Qu", m, N), ("Qxx_tmp", n * n, N), ("Quu_tmp", m * m, N),
                                    ("Qux_tmp", m * n, N), ("Qx_tmp", n, N), ("Qu_tmp", m, N)]:
                o[name] = tab([np.zeros(size) for _ in range(cnt)])
            nx = (C.c_int * (N + 1))(*([n] * (N + 1))); nu = (C.c_int * N)(*([m] * N))
            odV = np.zeros(2)
            T = lambda k: o[k][0]
            res = OL.oracle_tvlqr_backward_pass(nx, nu, N, tA, tB, tf, tQ, tR, tH, tq, tr, C.c_double(0.0),
                                                T("K"), T("d"), T("P"), T("p"), odV.ctypes.data_as(dp),
                                                T("Qxx"), T("Quu"), T("Qux"), T("Qx"), T("Qu"), T("Qxx_tmp"),
                                                T("Quu_tmp"), T("Qux_tmp"), T("Qx_tmp"), T("Qu_tmp"), 0, 0)
            assert res == -1
            for k in range(N):
                assert np.allclose(K[b, k].reshape(-1), o["K"][1][k], rtol=1e-9, atol=1e-10)
                assert np.allclose(d[b, k], o["d"][1][k], rtol=1e-9, atol=1e-10)
            assert np.allclose(P[b, 0].reshape(-1), o["P"][1][0], rtol=1e-9, atol=1e-9)
            assert np.allclose(dV[b], odV, rtol=1e-9)
        # closed-loop consistency of the forward pass: x+ = f + A x + B u, u = d - K x
        b = 0
        for k in range(N):
            Kk = K[b, k].T      # stored column-major m x n
            u = d[b, k] - Kk @ X[b, k]
            assert np.allclose(u, U[b, k], rtol=1e-10, atol=1e-10)
            assert np.allclose(f[b, k] + A[b, k] @ X[b, k] + Bm[b, k] @ u, X[b, k + 1], rtol=1e-10, atol=1e-10)


def test_tvlqr_workspace_diag_status_and_timing(oracle):
    """The device-resident TVLQR workspace (section B): diagonal-cost mode against the oracle, the
    per-problem Cholesky status (tvlqr.cpp:162-164), repeated launches without re-upload, and the
    CUDA-event timing entry point behind the kernel's roofline line."""
    L = altro_b200.load_library()
    L.altro_b200_tvlqr_ws_create.restype = C.c_void_p
    L.altro_b200_tvlqr_ws_create.argtypes = [C.c_int] * 4 + [C.c_bool, C.c_int]
    vp = C.c_void_p
    L.altro_b200_tvlqr_ws_upload.argtypes = [vp] + [dp] * 8
    L.altro_b200_tvlqr_ws_backward.argtypes = [vp, C.c_double]
    L.altro_b200_tvlqr_ws_download.argtypes = [vp, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
    L.altro_b200_tvlqr_ws_time_backward.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_float)]
    L.altro_b200_tvlqr_ws_bytes_per_knot.restype = C.c_long
    L.altro_b200_tvlqr_ws_bytes_per_knot.argtypes = [vp]
    L.altro_b200_tvlqr_ws_destroy.argtypes = [vp]
    assert not L.altro_b200_tvlqr_ws_create(8, 7, 3, 10, True, 0)        # shape not compiled in
    rng = np.random.default_rng(9)
    n, m, N, B = 6, 2, 40, 100
    A = np.eye(n) + 0.05 * rng.normal(size=(B, N, n, n))
    Bm = rng.normal(size=(B, N, n, m))
    f = 0.1 * rng.normal(size=(B, N, n))
    Qd = 0.5 + rng.uniform(size=(B, N + 1, n))
    Rd = 0.1 + rng.uniform(size=(B, N, m))
    Rd[7, 11] = -1e6                                                      # problem 7 fails at knot 11
    q, r = rng.normal(size=(B, N + 1, n)), rng.normal(size=(B, N, m))
    cm = lambda M: np.ascontiguousarray(np.swapaxes(M, -1, -2))
    Ac, Bc = cm(A), cm(Bm)
    ptr = lambda a: a.ctypes.data_as(dp)
    w = vp(L.altro_b200_tvlqr_ws_create(B, n, m, N, True, 0))
    assert w.value
    assert L.altro_b200_tvlqr_ws_bytes_per_knot(w) == 8 * (n * n + n * m + n + n + m + n + m + m * n + m + n * n + n)
    assert L.altro_b200_tvlqr_ws_upload(w, ptr(Ac), ptr(Bc), ptr(f), ptr(Qd), ptr(Rd), None, ptr(q), ptr(r)) == 0
    K = np.zeros((B, N, n, m)); d = np.zeros((B, N, m)); P = np.zeros((B, N + 1, n, n))
    p = np.zeros((B, N + 1, n)); dV = np.zeros((B, 2)); st = np.zeros(B, dtype=np.int32)
    for rep in range(2):                                                  # second launch: no upload
        assert L.altro_b200_tvlqr_ws_backward(w, 0.0) == 0
        assert L.altro_b200_tvlqr_ws_download(w, ptr(K), ptr(d), ptr(P), ptr(p), ptr(dV),
                                              st.ctypes.data_as(C.POINTER(C.c_int))) == 0
        assert st[7] == 11 and np.all(np.delete(st, 7) == -1)
    ms = C.c_float()
    assert L.altro_b200_tvlqr_ws_time_backward(w, 0.0, 5, C.byref(ms)) == 0 and ms.value > 0
    L.altro_b200_tvlqr_ws_destroy(w)
    OL = oracle.lib()
    for b in (0, 50, 99):
        def tab(arrs):
            keep = [np.ascontiguousarray(a, dtype=float).reshape(-1) for a in arrs]
            return (dp * len(keep))(*[a.ctypes.data_as(dp) for a in keep]), keep
        tA, k1 = tab(list(Ac[b])); tB, k2 = tab(list(Bc[b])); tf, k3 = tab(list(f[b]))
        tQ, k4 = tab(list(Qd[b])); tR, k5 = tab(list(Rd[b])); tH, k6 = tab([np.zeros(m * n)] * N)
        tq, k7 = tab(list(q[b])); tr, k8 = tab(list(r[b]))
        o = {}
        for name, size, cnt in [("K", m * n, N), ("d", m, N), ("P", n * n, N + 1), ("p", n, N + 1),
                                ("Qxx", n * n, N), ("Quu", m * m, N), ("Qux", m * n, N), ("Qx", n, N),
                                ("Qu", m, N), ("Qxx_tmp", n * n, N), ("Quu_tmp", m * m, N),
                                ("Qux_tmp", m * n, N), ("Qx_tmp", n, N), ("Qu_tmp", m, N)]:
            o[name] = tab([np.zeros(size) for _ in range(cnt)])
        nx = (C.c_int * (N + 1))(*([n] * (N + 1))); nu = (C.c_int * N)(*([m] * N))
        odV = np.zeros(2)
        T = lambda k: o[k][0]
        res = OL.oracle_tvlqr_backward_pass(nx, nu, N, tA, tB, tf, tQ, tR, tH, tq, tr, C.c_double(0.0),
                                            T("K"), T("d"), T("P"), T("p"), odV.ctypes.data_as(dp),
                                            T("Qxx"), T("Quu"), T("Qux"), T("Qx"), T("Qu"), T("Qxx_tmp"),
                                            T("Quu_tmp"), T("Qux_tmp"), T("Qx_tmp"), T("Qu_tmp"), 0, 1)
        assert res == -1
        for k in range(N):
            assert np.allclose(K[b, k].reshape(-1), o["K"][1][k], rtol=1e-9, atol=1e-10)
            assert np.allclose(d[b, k], o["d"][1][k], rtol=1e-9, atol=1e-10)
        assert np.allclose(P[b, 0].reshape(-1), o["P"][1][0], rtol=1e-9, atol=1e-9)
        assert np.allclose(p[b, 0], o["p"][1][0], rtol=1e-9, atol=1e-9)
        assert np.allclose(dV[b], odV, rtol=1e-9)


CONSUMER = r'''
// The reference's own end-to-end tests re-expressed against the facade
// (test/double_integrator_test.cpp:169-256, :258-375; test/pendulum_test.cpp:45-115).
#include "altro/altro_solver.hpp"
#include <cmath>
#include <cstdio>
#include <vector>
using namespace altro;
#define CHECK(cond) do { if (!(cond)) { std::printf("CHECK FAILED line %d: %s\n", __LINE__, #cond); return 1; } } while (0)

int double_integrator(bool control_bounds) {
  const int N = 10, n = 4, m = 2;
  const float h = 5.0f / N;
  std::vector<double> Q(n, 1.0), R(m, 1e-2), x0 = {control_bounds ? 2.0 : 1.0, 2.0, 0.0, 0.0}, xf(n, 0.0), uf(m, 0.0);
  ALTROSolver solver(N);
  CHECK(solver.SetDimension(n, m, 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetTimeStep(h, 0, LastIndex) == ErrorCodes::NoError);
  b200::DeviceDynamics model(b200::DeviceDynamics::DoubleIntegrator, {2});
  CHECK(solver.SetExplicitDynamics(model.Function(), model.Jacobian(), 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetLQRCost(n, m, Q.data(), R.data(), xf.data(), uf.data(), 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetInitialState(x0.data(), n) == ErrorCodes::NoError);
  auto goal = b200::DeviceConstraint::Goal(xf);
  CHECK(solver.SetConstraint(goal.Function(), goal.Jacobian(), n, ConstraintType::EQUALITY, "Goal constraint", N, 0, nullptr) == ErrorCodes::NoError);
  if (control_bounds) {
    auto box = b200::DeviceConstraint::InputBox(n, {1.0, 1.0});
    CHECK(solver.SetConstraint(box.Function(), box.Jacobian(), 2 * m, ConstraintType::INEQUALITY, "Control bounds", 0, N, nullptr) == ErrorCodes::NoError);
  }
  CHECK(solver.Initialize() == ErrorCodes::NoError);
  CHECK(solver.IsInitialized());
  std::vector<double> uinit(m, 0.0);
  CHECK(solver.SetState(x0.data(), n, 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetInput(uinit.data(), m, 0, LastIndex) == ErrorCodes::NoError);
  AltroOptions opts;
  opts.penalty_scaling = 100;
  if (control_bounds) opts.penalty_initial = 100;
  solver.SetOptions(opts);
  CHECK(solver.Solve() == SolveStatus::Success);
  std::vector<double> xN(n), u0(m);
  CHECK(solver.GetState(xN.data(), N) == ErrorCodes::NoError);
  double dist = 0; for (double v : xN) dist += v * v;
  CHECK(std::sqrt(dist) < 1e-4);
  CHECK(solver.GetIterations() == (control_bounds ? 5 : 3));     // :255, :374
  if (control_bounds) { solver.GetInput(u0.data(), 0); CHECK(std::fabs(u0[0] + 1.0) < 1e-4 && std::fabs(u0[1] + 1.0) < 1e-4); }
  CHECK(solver.GetStatus() == SolveStatus::Success);
  std::vector<double> K(m * n);
  CHECK(solver.GetFeedbackGain(K.data(), 0) == ErrorCodes::NoError);
  // KnotPointData-named views (what the reference's tests read through solver_->data_[k])
  std::vector<double> Kf(m * n), A(n * n), Bk(n * m);
  CHECK(solver.GetKnotPointField("K", Kf.data(), 0) == ErrorCodes::NoError);
  for (int i = 0; i < m * n; ++i) CHECK(Kf[i] == K[i]);
  CHECK(solver.GetKnotPointField("A", A.data(), 0) == ErrorCodes::NoError);
  CHECK(solver.GetKnotPointField("B", Bk.data(), 0) == ErrorCodes::NoError);
  CHECK(A[0] == 1.0 && std::fabs(A[0 + n * 2] - h) < 1e-7 && std::fabs(Bk[2 + n * 0] - h) < 1e-7);  // test_utils.cpp:18-41
  CHECK(solver.GetKnotPointField("nope", A.data(), 0) == ErrorCodes::BadIndex);
  // error conventions
  CHECK(solver.SetDimension(n, m) == ErrorCodes::SolverAlreadyInitialized);
  CHECK(solver.GetState(xN.data(), N + 5) == ErrorCodes::BadIndex);
  return 0;
}

// The declared-but-undefined bound setters of the reference (altro_solver.hpp:257-290) as
// INEQUALITY rows: same control-box problem as double_integrator(true), same pinned answers
// (double_integrator_test.cpp:367-374), plus the dual getters / setters (:359, :416).
int double_integrator_bound_api() {
  const int N = 10, n = 4, m = 2;
  const float h = 5.0f / N;
  std::vector<double> Q(n, 1.0), R(m, 1e-2), x0 = {2.0, 2.0, 0.0, 0.0}, xf(n, 0.0), uf(m, 0.0);
  ALTROSolver solver(N);
  CHECK(solver.SetDimension(n, m, 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetTimeStep(h, 0, LastIndex) == ErrorCodes::NoError);
  b200::DeviceDynamics model(b200::DeviceDynamics::DoubleIntegrator, {2});
  CHECK(solver.SetExplicitDynamics(model.Function(), model.Jacobian(), 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetLQRCost(n, m, Q.data(), R.data(), xf.data(), uf.data(), 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetInitialState(x0.data(), n) == ErrorCodes::NoError);
  auto goal = b200::DeviceConstraint::Goal(xf);
  std::vector<ConstraintIndex> goal_idx;
  CHECK(solver.SetConstraint(goal.Function(), goal.Jacobian(), n, ConstraintType::EQUALITY, "Goal constraint", N, 0, &goal_idx) == ErrorCodes::NoError);
  CHECK(goal_idx.size() == 1 && goal_idx[0].KnotPointIndex() == N);
  std::vector<double> umax = {1.0, 1.0}, umin = {-1.0, -1.0};
  CHECK(solver.SetInputUpperBound(umax.data(), 0, N) == ErrorCodes::NoError);
  CHECK(solver.SetInputLowerBound(umin.data(), 0, N) == ErrorCodes::NoError);
  CHECK(solver.SetInputUpperBound(umax.data(), N) == ErrorCodes::InvalidOptAtTerminalKnotPoint);
  CHECK(solver.Initialize() == ErrorCodes::NoError);
  std::vector<double> uinit(m, 0.0);
  CHECK(solver.SetInput(uinit.data(), m, 0, LastIndex) == ErrorCodes::NoError);
  AltroOptions opts;
  opts.penalty_scaling = 100;
  opts.penalty_initial = 100;
  solver.SetOptions(opts);
  CHECK(solver.Solve() == SolveStatus::Success);
  CHECK(solver.GetIterations() == 5);
  std::vector<double> u0(m), xN(n);
  solver.GetInput(u0.data(), 0);
  CHECK(std::fabs(u0[0] + 1.0) < 1e-4 && std::fabs(u0[1] + 1.0) < 1e-4);
  solver.GetState(xN.data(), N);
  double dist = 0; for (double v : xN) dist += v * v;
  CHECK(std::sqrt(dist) < 1e-4);
  // goal dual: readable, writable, and what the KnotPointData view shows
  std::vector<double> z(n), zview(n + 0), z2(n);
  CHECK(solver.GetDualGeneral(z.data(), goal_idx[0]) == ErrorCodes::NoError);
  double zn = 0; for (double v : z) zn += v * v;
  CHECK(zn > 0);
  std::vector<double> zall(n + 2 * m);   // all constraint rows of the handle: goal + two bound slots
  CHECK(solver.GetKnotPointField("z_", zall.data(), N) == ErrorCodes::NoError);
  for (int i = 0; i < n; ++i) CHECK(zall[i] == z[i]);
  for (int i = 0; i < n; ++i) z2[i] = 0.5 * z[i];
  CHECK(solver.SetDualGeneric(z2.data(), goal_idx[0]) == ErrorCodes::NoError);
  CHECK(solver.GetDualGeneral(z.data(), goal_idx[0]) == ErrorCodes::NoError);
  for (int i = 0; i < n; ++i) CHECK(z[i] == z2[i]);
  // the lower input bound is active at knot 0: its dual is negative (negative orthant)
  std::vector<double> zk(n + 2 * m);
  CHECK(solver.GetKnotPointField("z", zk.data(), 0) == ErrorCodes::NoError);
  CHECK(zk[n + m] < 0 && zk[n + m + 1] < 0 && zk[n] == 0 && zk[n + 1] == 0);
  return 0;
}

int pendulum() {
  const int n = 2, m = 1, N = 50;
  const float h = 3.0f / N;
  std::vector<double> Qd(n, 1e-2), Rd(m, 1e-3), Qdf(n, 1.0), x0(n, 0.0), xf = {M_PI, 0.0}, uf(m, 0.0);
  ALTROSolver solver(N);
  CHECK(solver.SetDimension(n, m, 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetTimeStep(h, 0, LastIndex) == ErrorCodes::NoError);
  b200::DeviceDynamics model(b200::DeviceDynamics::Pendulum);
  CHECK(solver.SetExplicitDynamics(model.Function(), model.Jacobian(), 0, LastIndex) == ErrorCodes::NoError);
  CHECK(solver.SetLQRCost(n, m, Qd.data(), Rd.data(), xf.data(), uf.data(), 0, N) == ErrorCodes::NoError);
  CHECK(solver.SetLQRCost(n, m, Qdf.data(), Rd.data(), xf.data(), uf.data(), N) == ErrorCodes::NoError);
  CHECK(solver.SetInitialState(x0.data(), n) == ErrorCodes::NoError);
  CHECK(solver.Initialize() == ErrorCodes::NoError);
  std::vector<double> u0(m, 0.1);
  CHECK(solver.SetInput(u0.data(), m, 0, LastIndex) == ErrorCodes::NoError);
  AltroOptions opts; opts.iterations_max = 20; solver.SetOptions(opts);
  CHECK(solver.Solve() == SolveStatus::Success);
  std::vector<double> xN(n);
  solver.GetState(xN.data(), N);
  const double e0 = xN[0] - 3.12099917161669, e1 = xN[1] - 0.0011966258762942175;   // pendulum_test.cpp:111
  CHECK(std::sqrt(e0 * e0 + e1 * e1) < 1e-5);
  CHECK(solver.GetIterations() <= 10);
  // a host lambda must be rejected, not run on the CPU
  ALTROSolver s2(N);
  s2.SetDimension(n, m);
  auto dyn = [](double*, const double*, const double*, float) {};
  CHECK(s2.SetExplicitDynamics(dyn, dyn) == ErrorCodes::DynamicsFunNotSet);
  return 0;
}

int main() {
  if (double_integrator(false)) return 1;
  if (double_integrator(true)) return 2;
  if (pendulum()) return 3;
  if (double_integrator_bound_api()) return 4;
  std::printf("FACADE OK\n");
  return 0;
}
'''


def test_cpp_facade_reference_tests(tmp_path):
    src = tmp_path / "facade_tests.cpp"
    src.write_text(CONSUMER)
    exe = tmp_path / "facade_tests"
    libdir = os.path.join(ROOT, "altro_b200")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                        str(exe), "-L", libdir, "-laltro_b200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "FACADE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_open_loop_rollout_and_calc_cost(oracle):
    from altro_b200 import problems as PR
    P = PR.scotty(B=16, N=30, n=4)
    s = altro_b200.make_solver(P)
    s.OpenLoopRollout()
    X = s.GetStates()
    cost = s.CalcCost()
    for b in (0, 7, 15):
        import ctypes as C
        sp, keep = oracle.make_batch_spec(P)
        h = C.c_void_p(oracle.lib().oracle_batch_make_solver(C.byref(sp), b))
        so = oracle.OracleSolver.__new__(oracle.OracleSolver)
        so.L, so.N, so.n, so.m, so.h, so._keep = oracle.lib(), P.N, P.n, P.m, h, []
        so.OpenLoopRollout()
        assert np.abs(so.states() - X[b]).max() < 1e-12
        assert abs(so.CalcCost() - cost[b]) <= 1e-12 * max(1.0, abs(cost[b]))
    s.close()
