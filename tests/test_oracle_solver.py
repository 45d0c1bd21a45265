"""Oracle solver loop vs the reference's end-to-end goldens:
solver_impl_test.cpp, alilqr_test.cpp, double_integrator_test.cpp, pendulum_test.cpp,
bicycle_test.cpp and the 200-step MPC golden test/scotty_mpc.json (tests/golden/)."""
import json
import math
import os

import numpy as np
import pytest

from altro_b200 import problems as PR
from test_oracle_cones_tvlqr import (D0_EXPECTED, K0_EXPECTED, XN_EXPECTED, YN_EXPECTED,
                                     tvlqr_problem)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EQ, ID, INEQ, SOC = 0, 1, 2, 3


def di_lq_solver(O):
    """SolverImplTest::InitializeDoubleIntegratorSolver (solver_impl_test.cpp:19-92)."""
    pr = tvlqr_problem(np.float32(0.01))
    n, m, N = pr["n"], pr["m"], pr["N"]
    s = O.OracleSolver(N, n, m)
    s.SetTimeStep(0.01)
    s.SetInitialState(pr["x0"])
    for k in range(N):
        s.SetLinearDynamics(k, pr["A"], pr["B"], pr["f"])
        s.SetDiagonalCost(k, pr["Qd"], pr["Rd"], pr["q"], pr["r"], 0.0)
    s.SetDiagonalCost(N, pr["Qfd"], None, pr["q"], None, 0.0)
    assert s.Initialize() == 0
    return s, pr


def test_solverimpl_tvlqr(oracle):
    """solver_impl_test.cpp:110-155"""
    s, pr = di_lq_solver(oracle)
    n, m, N = pr["n"], pr["m"], pr["N"]
    s.BackwardPass(); s.BackwardPass()
    assert np.linalg.norm(s.get(0, "K_", (m, n)) - K0_EXPECTED) < 1e-6
    assert np.linalg.norm(s.get(0, "d_") - D0_EXPECTED) < 1e-6
    s.LinearRollout()
    assert np.abs(s.get(N, "x_") - XN_EXPECTED).max() < 1e-6
    assert np.abs(s.get(N, "y_") - YN_EXPECTED).max() < 1e-5
    s.CalcCostGradient()
    assert s.Stationarity() < 1e-10


def _compute_gains(s, N):
    for k in range(N + 1):
        for op in ("CalcCostGradient", "CalcConstraints", "CalcConstraintJacobians",
                   "CalcProjectedDuals", "CalcConicJacobians", "CalcDynamicsExpansion"):
            s.KnotOp(k, op)
    s.CalcExpansions()
    s.BackwardPass()


def test_merit_function_goldens(oracle):
    """solver_impl_test.cpp:188-266: phi(1), phi'(1), phi(0), phi'(0) rel 1e-6 + finite diff."""
    s, pr = di_lq_solver(oracle)
    n, m, N, x0 = pr["n"], pr["m"], pr["N"], pr["x0"]
    xf = np.array([-1.0, 2, 0, 0])
    for k in range(N):
        theta = k / N
        s.set(k, "x_", x0 + (xf - x0) * theta)
        s.set(k, "u_", np.full(m, theta))
    s.set(N, "x_", xf)
    s.CopyTrajectory()
    _compute_gains(s, N)
    phi, dphi = s.MeritFunction(1.0)
    assert abs(phi - 25992.822836536347) / 25992.822836536347 < 1e-6
    assert abs(dphi - -43.52330058003784) / 43.52330058003784 < 1e-6
    eps = 1e-6
    phi1 = s.MeritFunction(1.0 + eps, want_derivative=False)
    assert abs(dphi - (phi1 - phi) / eps) / abs(dphi) < 1e-6
    phi, dphi = s.MeritFunction(0.0)
    assert abs(phi - 26039.092492842017) / 26039.092492842017 < 1e-6
    assert abs(dphi - -49.01601203132092) / 49.01601203132092 < 1e-6


def test_forward_pass_lq_alpha_one(oracle):
    """solver_impl_test.cpp:268-315"""
    s, pr = di_lq_solver(oracle)
    N, m = pr["N"], pr["m"]
    for k in range(N):
        s.set(k, "u_", np.full(m, k / N))
    s.OpenLoopRollout()
    s.CopyTrajectory()
    _compute_gains(s, N)
    phi, dphi = s.MeritFunction(1.0)
    assert abs(dphi) < 1e-8
    err, alpha = s.ForwardPass()
    assert alpha == 1.0


def test_alilqr_staged(oracle):
    """alilqr_test.cpp:107-215 (pendulum N=20 goal-constrained, manual AL loop, c2=0.1)."""
    O = oracle
    N, n, m = 20, 2, 1
    h = np.float32(np.float32(2.0) / float(N))
    xf = np.array([math.pi, 0.0])
    Qd, Rd, Qdf = np.full(n, 1e-2), np.full(m, 1e-3), np.full(n, 1.0)
    s = O.OracleSolver(N, n, m)
    s.SetTimeStep(h)
    s.SetModel(O.MODEL_PENDULUM)
    for k in range(N):
        s.SetLQRCost(k, Qd, Rd, xf, np.zeros(m))
    s.SetLQRCost(N, Qdf, Rd, xf, np.zeros(m))
    s.AddSelectorConstraint(N, EQ, [0, 1], [-1.0, -1.0], list(xf))
    s.SetInitialState(np.zeros(n))
    assert s.Initialize() == 0
    s.SetOptions(O.default_options(ls_c1=1e-4, ls_c2=0.1))
    s.SetInput(np.full(m, 0.1))
    s.OpenLoopRollout()
    s.CopyTrajectory()
    phi0 = s.MeritFunction(0.0, want_derivative=False)
    assert phi0 == pytest.approx(10.632455092693577, abs=1e-3)

    def refresh():
        for k in range(N + 1):
            for op in ("CalcDynamicsExpansion", "CalcConstraints", "CalcConstraintJacobians",
                       "CalcProjectedDuals", "CalcConicJacobians", "CalcCostGradient"):
                s.KnotOp(k, op)

    def stage(allow_small_grad):
        dist = None
        for it in range(6):
            s.CalcExpansions(); s.BackwardPass()
            err, alpha = s.ForwardPass()
            if allow_small_grad and err == 21:
                break
            assert err == 0
            s.Stationarity()
            dist = np.linalg.norm(s.get(N, "x_") - xf)
            s.CopyTrajectory()
        return dist

    refresh()
    d0 = stage(False)
    assert d0 == pytest.approx(0.04186387, abs=1e-3)
    s.DualUpdate(); s.PenaltyUpdate(); refresh()
    d1 = stage(True)
    assert d1 < d0 / 5
    s.DualUpdate(); s.PenaltyUpdate(); s.PenaltyUpdate(); refresh()
    d2 = stage(True)
    assert d2 < 1e-4


@pytest.mark.parametrize("variant,iters", [("goal", 3), ("ubox", 5), ("usoc", 9)])
def test_double_integrator_iteration_counts(oracle, variant, iters):
    """double_integrator_test.cpp:255, 374, 491 (asserted ==) + terminal tolerances."""
    P = PR.double_integrator(N=10, variant=variant)
    r = oracle.solve_batch(P, nthreads=1)
    assert r["status"][0] == 0 and r["iters"][0] == iters
    assert np.linalg.norm(r["X"][0, -1]) < 1e-4
    if variant == "ubox":
        assert np.allclose(r["U"][0, 0], -1.0, atol=1e-4)
    if variant == "usoc":
        assert abs(np.linalg.norm(r["U"][0, 0]) - 1.0) < 1e-2


def test_double_integrator_unconstrained(oracle):
    """double_integrator_test.cpp:66-167: Success within iterations_max=3."""
    P = PR.double_integrator(N=10, variant="unconstrained")
    r = oracle.solve_batch(P, nthreads=1)
    d = np.linalg.norm(r["X"][0, -1])
    assert r["status"][0] == 0 and 1e-3 < d < np.linalg.norm(P.x0[0])


def test_double_integrator_dynamics_golden(oracle):
    """double_integrator_test.cpp:36-64"""
    xn = oracle.model_dynamics(oracle.MODEL_DI, [2], [0.1, 0.2, 0.3, 0.4], [10.1, -20.4], 0.01)
    exp = [0.10350500000000001, 0.20298000000000002, 0.40099999999999997, 0.19600000000000004]
    assert np.linalg.norm(xn - exp) < 1e-8
    J = oracle.model_jacobian(oracle.MODEL_DI, [2], [0.1, 0.2, 0.3, 0.4], [10.1, -20.4], 0.01)
    h = float(np.float32(0.01)); b = float(np.float32(np.float32(0.01) ** 2 / 2))
    Je = np.array([[1, 0, h, 0, b, 0], [0, 1, 0, h, 0, b], [0, 0, 1, 0, h, 0], [0, 0, 0, 1, 0, h]])
    assert np.linalg.norm(J - Je) < 1e-8


def test_pendulum_dynamics_goldens(oracle):
    """pendulum_test.cpp:14-43"""
    O = oracle
    xn = O.model_dynamics(O.MODEL_PENDULUM, [], [0.1, -0.4], [1.34], 0.05)
    assert np.linalg.norm(xn - [0.08445158545673655, -0.21395149094594346]) < 1e-6
    J = O.model_jacobian(O.MODEL_PENDULUM, [], [0.1, -0.4], [1.34], 0.05)
    Je = np.array([[0.9755975228465564, 0.0495, 0.005000000000000001],
                   [-0.967268640223389, 0.9557742592228808, 0.198]])
    assert np.linalg.norm(J - Je) < 1e-6


def test_bicycle_dynamics_goldens(oracle):
    """bicycle_test.cpp:27-51"""
    O = oracle
    x = [1, 0.5, 15 * math.pi / 180, -5 * math.pi / 180]; u = [1.1, 0.2]
    xd = O.model_continuous(O.MODEL_BICYCLE4, [2.7, 1.5], x, u)
    assert np.linalg.norm(xd - [1.0750584102061864, 0.23291503739549996, -0.03560171424038893, 0.2]) < 1e-10
    J = O.model_continuous_jacobian(O.MODEL_BICYCLE4, [2.7, 1.5], x, u)
    Je = np.array([-0.0, -0.0, -0.23291503739549996, -0.1290938153359409, 0.9773258274601694, 0.0, 0.0,
                   0.0, 1.0750584102061864, 0.5958541510862063, 0.21174094308681812, 0.0, 0.0, 0.0, 0.0,
                   0.409087550891862, -0.03236519476398994, -0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]).reshape(4, 6)
    assert np.linalg.norm(J - Je) < 1e-6


@pytest.mark.parametrize("model,params,n,m", [(4, [2.7, 1.5], 5, 2), (5, [6, 2], 6, 2),
                                               (5, [12, 4], 12, 4), (3, [2.7, 1.5], 4, 2)])
def test_model_jacobians_finite_difference(oracle, model, params, n, m):
    """The models that do not exist in the reference (bicycle5, chain): discrete Jacobian vs FD."""
    O = oracle
    rng = np.random.default_rng(5)
    x, u, h = rng.uniform(-0.4, 0.4, n), rng.uniform(-0.5, 0.5, m), 0.03
    if model in (3, 4):
        u[0] += 1.0
        if n == 5:
            x[4] += 1.0
    J = O.model_jacobian(model, params, x, u, h)
    Jfd = np.zeros_like(J)
    eps = 1e-6
    for j in range(n + m):
        dx, du = np.zeros(n), np.zeros(m)
        if j < n:
            dx[j] = eps
        else:
            du[j - n] = eps
        Jfd[:, j] = (O.model_dynamics(model, params, x + dx, u + du, h)
                     - O.model_dynamics(model, params, x - dx, u - du, h)) / (2 * eps)
    assert np.abs(J - Jfd).max() < 1e-7


def test_pendulum_unconstrained(oracle):
    """pendulum_test.cpp:45-115: Success, x_N golden +-1e-5, <= 10 iterations."""
    P = PR.pendulum(B=1, N=50, tf=3.0, perturb=False, iterations_max=20)
    r = oracle.solve_batch(P, nthreads=1)
    assert r["status"][0] == 0 and r["iters"][0] <= 10
    assert np.linalg.norm(r["X"][0, -1] - [3.12099917161669, 0.0011966258762942175]) < 1e-5
    assert (r["iters"][0], r["merit_evals"][0]) == (10, 24)


def test_pendulum_goal_constrained(oracle):
    """pendulum_test.cpp:117-203"""
    P = PR.pendulum(B=1, N=20, tf=2.0, perturb=False, goal_constraint=True, iterations_max=100)
    r = oracle.solve_batch(P, nthreads=1)
    assert r["status"][0] == 0 and r["iters"][0] <= 10
    assert np.linalg.norm(r["X"][0, -1] - [math.pi, 0]) < 1e-4
    assert (r["iters"][0], r["merit_evals"][0]) == (9, 25)


def test_bicycle_turn90(oracle):
    """bicycle_test.cpp:53-138"""
    P = PR.bicycle_turn90()
    r = oracle.solve_batch(P, nthreads=1)
    assert np.linalg.norm(r["X"][0, -1] - P.xref[0]) < 1e-2
    assert (r["iters"][0], r["merit_evals"][0]) == (17, 40)


def scotty_mpc_run(O, Nsim=200):
    """BicycleMPC fixture + TrackingMPC_2Solves (bicycle_test.cpp:140-337)."""
    xref, uref, h = PR.load_scotty()
    n, m, N = 4, 2, 30
    Qd, Rd = np.full(n, 1e-2), np.full(m, 1e-3)
    s = O.OracleSolver(N, n, m)
    s.SetModel(O.MODEL_BICYCLE4, [2.7, 1.5])
    s.SetTimeStep(h)
    for k in range(N + 1):
        s.SetLQRCost(k, Qd, Rd, xref[k], uref[k])
    dmax = 60 * math.pi / 180.0
    for k in range(N + 1):
        s.AddSelectorConstraint(k, INEQ, [3, 3], [1.0, -1.0], [-dmax, -dmax])
    s.SetInitialState(xref[0])
    assert s.Initialize() == 0
    u0 = np.array([uref[0][0], 0.0])
    s.SetInput(u0)
    for k in range(N + 1):
        s.SetState(xref[k], k)
    s.SetOptions(O.default_options(iterations_max=80, use_backtracking_linesearch=1))
    c_u = 0.5 * u0 @ (Rd * u0)
    x_sim = [xref[0].copy()]
    u_sim, iters, errs, evals = [], [], [], 0
    for it in range(Nsim):
        status = s.Solve()
        assert status == 0
        iters.append(s.GetIterations())
        evals += s.GetMeritEvals()
        u = s.GetInput(0)
        xn = O.model_dynamics(O.MODEL_BICYCLE4, [2.7, 1.5], x_sim[-1], u, h)
        u_sim.append(u); x_sim.append(xn)
        errs.append(np.linalg.norm(xn - xref[it + 1]))
        for k in range(N + 1):
            xk = xref[k + it + 1]
            q = -(Qd * xk)
            c = -(0.5 * q @ xk)
            if k < N:
                c += c_u
            s.UpdateLinearCosts(k, q, None, c)
        s.SetInitialState(xn)
        s.ShiftTrajectory()
    return np.array(x_sim), np.array(u_sim), iters, errs, evals


def test_scotty_mpc_golden(oracle):
    """All 200 warm-started MPC solves of test/scotty_mpc.json: identical iteration counts,
    closed-loop states/inputs to ~1e-12."""
    gold = json.load(open(os.path.join(GOLDEN, "scotty_mpc.json")))
    x_sim, u_sim, iters, errs, evals = scotty_mpc_run(oracle, 200)
    assert iters == gold["solve_iters"] and sum(iters) == 627
    assert np.abs(x_sim - np.array(gold["state_trajectory"])).max() < 1e-11
    assert np.abs(u_sim - np.array(gold["input_trajectory"])).max() < 1e-9
    assert np.abs(np.array(errs) - np.array(gold["tracking_error"])).max() < 1e-11
    assert evals == 1523


def test_batch_driver_matches_single(oracle):
    """oracle_batch_solve (OpenMP) == per-problem solves; REF_WINDOW and REF_GOAL paths."""
    P = PR.scotty(B=6, N=30, n=4)
    r8 = oracle.solve_batch(P, nthreads=4)
    r1 = oracle.solve_batch(P, nthreads=1)
    assert np.array_equal(r8["X"], r1["X"]) and np.array_equal(r8["iters"], r1["iters"])
    sub = oracle.solve_batch(P, b0=2, b1=5, nthreads=2)
    assert np.array_equal(sub["X"], r1["X"][2:5])
    assert np.all(r1["status"] == 0)
