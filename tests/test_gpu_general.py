"""General problem data on the device, checked against the oracle
through the reference's own callback interface:

  * dense quadratic cost Q, R and the u'Hx cross term (knotpoint_data.cpp:616-708);
  * general affine constraints c = J [x;u] + e with a dense Jacobian, and a nonlinear constraint
    family (keep-out disc), any cone (knotpoint_data.cpp:473-613);
  * the reference's knot-point goldens for the AL cost / gradient / Hessian
    (src/altro/solver/test/knotpoint_data_test.cpp:233-524: INEQUALITY, EQUALITY, SOC out of /
    below / inside the cone at x = (2,2,2), u = (10,10), rho = 1.2) reproduced ON THE GPU through
    the KnotPointData views;
  * KnotPointData constraint / dual / Hessian views and the dual getters / setters;
  * the declared-but-undefined bound setters (altro_solver.hpp:257-290) as INEQUALITY rows.
"""
import numpy as np
import pytest

import altro_b200
from altro_b200 import problems as PR
from altro_b200.solver import BatchSolver, default_options

pytestmark = pytest.mark.gpu

EQ, ID, INEQ, SOC = 0, 1, 2, 3


def di_solver(B, N=10, tf=5.0, x0=None):
    """Double integrator (test/double_integrator_test.cpp:60-105), B problems."""
    s = BatchSolver(N, B)
    s.SetDimension(4, 2)
    s.SetTimeStep(tf / N)
    s.SetExplicitDynamics(PR.MODEL_DOUBLE_INTEGRATOR, [2])
    x0 = np.tile([1.0, 2.0, 0.0, 0.0], (B, 1)) if x0 is None else x0
    s.SetInitialState(x0)
    return s, x0


def di_oracle(oracle, N, tf, x0):
    o = oracle.OracleSolver(N, 4, 2)
    o.SetTimeStep(tf / N)
    o.SetModel(oracle.MODEL_DI, [2])
    o.SetInitialState(x0)
    return o


def assert_same_solve(s, o, b, status_o, tol=1e-6):   # parity_util.STATE_TOL
    X, U = s.GetStates()[b], s.GetInputs()[b]
    assert s.GetStatus()[b] == status_o
    assert s.GetIterations()[b] == o.GetIterations()
    scale = max(1.0, np.abs(o.states()).max())
    assert np.abs(X - o.states()).max() <= tol * scale
    assert np.abs(U - o.inputs()).max() <= tol * max(1.0, np.abs(o.inputs()).max())
    assert abs(s.GetFinalObjective()[b] - o.GetFinalPhi()) <= 1e-7 * max(1.0, abs(o.GetFinalPhi()))


def spd(rng, k, lo):
    M = rng.normal(size=(k, k))
    return M @ M.T / k + lo * np.eye(k)


def test_dense_quadratic_cost_matches_oracle(oracle):
    """SetQuadraticCost with dense Q, R and the cross term H (knotpoint_data.cpp:616-708) on the
    nonlinear pendulum: ~11 iterations, every one through the dense cost / gradient / Hessian."""
    rng = np.random.default_rng(11)
    n, m, N, B = 2, 1, 30, 6
    Q = np.array([[1e-2, 4e-3], [4e-3, 2e-2]])
    R = np.array([[2e-3]])
    H = np.array([[1e-3, -2e-3]])
    Qf = np.array([[1.0, 0.2], [0.2, 1.5]])
    xf = np.array([np.pi, 0.0])
    q, r, c = -Q @ xf, -H @ xf, 0.5 * xf @ Q @ xf
    qf, cf = -Qf @ xf, 0.5 * xf @ Qf @ xf
    x0 = rng.uniform(-0.5, 0.5, size=(B, n))
    s = BatchSolver(N, B)
    s.SetDimension(n, m)
    s.SetTimeStep(3.0 / N)
    s.SetExplicitDynamics(PR.MODEL_PENDULUM, [])
    s.SetInitialState(x0)
    s.SetQuadraticCost(Q, R, H, q, r, c, 0, N)
    s.SetQuadraticCost(Qf, None, None, qf, None, cf, N, N + 1)
    s.Initialize()
    s.SetInput(np.array([0.1]))
    s.SetOptions(default_options(iterations_max=50))
    s.Solve()
    # the expansion views carry the dense blocks
    lux = s.GetField("lux")[0, 3].reshape(m, n, order="F")
    lxx = s.GetField("lxx")[0, 3].reshape(n, n, order="F")
    assert np.array_equal(lux, H) and np.array_equal(lxx, Q)
    assert np.array_equal(s.GetField("lxx")[0, N].reshape(n, n, order="F"), Qf)
    for b in range(B):
        o = oracle.OracleSolver(N, n, m)
        o.SetTimeStep(3.0 / N)
        o.SetModel(oracle.MODEL_PENDULUM, [])
        o.SetInitialState(x0[b])
        for k in range(N):
            o.SetQuadraticCost(k, Q, R, H, q, r, c)
        o.SetQuadraticCost(N, Qf, np.zeros((m, m)), np.zeros((m, n)), qf, np.zeros(m), cf)
        o.Initialize()
        o.SetInput(np.array([0.1]))
        o.SetOptions(oracle.default_options(iterations_max=50))
        st = o.Solve()
        assert st == 0 and o.GetIterations() > 5
        assert_same_solve(s, o, b, st)
    s.close()


def test_general_affine_constraints_match_oracle(oracle):
    """Dense-Jacobian affine rows on (x,u): a coupled input constraint u0 + u1 <= 0.6,
    0.5 x2 - u1 <= 1 (INEQUALITY) on every input knot plus the goal as a dense EQUALITY block."""
    N, tf, n, m, B = 10, 5.0, 4, 2, 4
    rng = np.random.default_rng(5)
    x0 = np.array([1.0, 2.0, 0.0, 0.0]) + 0.2 * rng.normal(size=(B, n))
    J = np.zeros((2, n + m))
    J[0, n + 0], J[0, n + 1] = 1.0, 1.0
    J[1, 2], J[1, n + 1] = 0.5, -1.0
    e = np.array([-0.6, -1.0])
    Jg = np.hstack([np.eye(n), np.zeros((n, m))])
    s, _ = di_solver(B, N, tf, x0)
    s.SetLQRCost(np.ones(n), np.full(m, 1e-2), np.zeros(n), np.zeros(m), 0, N + 1)
    s.SetConstraintAffine(INEQ, J, e, 0, N)
    s.SetConstraintAffine(EQ, Jg, np.zeros(n), N, N + 1)
    s.Initialize()
    s.SetInput(np.zeros(m))
    s.SetOptions(default_options(penalty_scaling=100.0, penalty_initial=10.0, iterations_max=40))
    s.Solve()
    U = s.GetInputs()
    assert (s.GetStatus() == 0).all()
    assert (U[:, :, 0] + U[:, :, 1] <= 0.6 + 1e-4).all()
    for b in range(B):
        o = di_oracle(oracle, N, tf, x0[b])
        for k in range(N + 1):
            o.SetLQRCost(k, np.ones(n), np.full(m, 1e-2), np.zeros(n), np.zeros(m))
        for k in range(N):
            o.AddCallbackConstraint(k, INEQ, 2, lambda x, u: J @ np.concatenate([x, u]) + e, lambda x, u: J)
        o.AddCallbackConstraint(N, EQ, n, lambda x, u: x.copy(), lambda x, u: Jg)
        o.Initialize()
        o.SetInput(np.zeros(m))
        o.SetOptions(oracle.default_options(penalty_scaling=100.0, penalty_initial=10.0, iterations_max=40))
        st = o.Solve()
        assert_same_solve(s, o, b, st, tol=1e-6)
    # duals through the KnotPointData view and the dual getter agree, and are active where the
    # coupled input bound binds (negative orthant: z <= 0)
    z = s.GetField("z")
    zk = s.GetDualGeneral(0, 0, 2)
    assert z.shape == (B, N + 1, 2 + n) and np.array_equal(z[:, 0, :2], zk)
    assert (zk <= 0).all() and (zk[:, 1] < -1e-3).all()
    s.close()


def test_nonlinear_disc_constraint_matches_oracle(oracle):
    """Keep-out disc (nonlinear constraint family) on the double integrator's position."""
    N, tf, n, m, B = 20, 5.0, 4, 2, 3
    x0 = np.array([[2.1, 2.0, 0.0, 0.0], [2.2, 1.9, 0.0, 0.0], [1.8, 2.1, 0.0, 0.0]])
    disc = np.array([1.0, 1.15, 0.5])

    def con(x, u):
        return np.array([disc[2] ** 2 - (x[0] - disc[0]) ** 2 - (x[1] - disc[1]) ** 2])

    def jac(x, u):
        Jm = np.zeros((1, n + m))
        Jm[0, 0], Jm[0, 1] = -2 * (x[0] - disc[0]), -2 * (x[1] - disc[1])
        return Jm

    s, _ = di_solver(B, N, tf, x0)
    s.SetLQRCost(np.ones(n), np.full(m, 1e-1), np.zeros(n), np.zeros(m), 0, N)
    s.SetLQRCost(np.full(n, 100.0), np.full(m, 1e-1), np.zeros(n), np.zeros(m), N, N + 1)
    s.SetConstraintDisc(0, 1, disc, 1, N + 1)
    s.Initialize()
    s.SetInput(np.zeros(m))
    opt = dict(penalty_scaling=10.0, penalty_initial=10.0, iterations_max=60)
    s.SetOptions(default_options(**opt))
    s.Solve()
    X = s.GetStates()
    ok = s.GetStatus() == 0
    assert ok.all()
    d2 = (X[:, 1:, 0] - disc[0]) ** 2 + (X[:, 1:, 1] - disc[1]) ** 2
    assert (d2 >= disc[2] ** 2 - 1e-3).all() and (d2.min(axis=1) < disc[2] ** 2 + 0.05).all()  # it binds
    cv = s.GetField("constraint_val")
    assert np.allclose(cv[:, 1:, 0], disc[2] ** 2 - d2, rtol=0, atol=1e-12)
    for b in range(B):
        o = di_oracle(oracle, N, tf, x0[b])
        for k in range(N):
            o.SetLQRCost(k, np.ones(n), np.full(m, 1e-1), np.zeros(n), np.zeros(m))
        o.SetLQRCost(N, np.full(n, 100.0), np.full(m, 1e-1), np.zeros(n), np.zeros(m))
        for k in range(1, N + 1):
            o.AddCallbackConstraint(k, INEQ, 1, con, jac)
        o.Initialize()
        o.SetInput(np.zeros(m))
        o.SetOptions(oracle.default_options(**opt))
        st = o.Solve()
        assert_same_solve(s, o, b, st, tol=1e-6)
    s.close()


# --------------------------------------------------------------------------------------------------
# src/altro/solver/test/knotpoint_data_test.cpp:136-524.  The fixture's constraint is nonlinear (two
# spheres and u0 + u1), but every golden is taken at ONE point, x = (2,2,2), u = (10,10), where the
# AL terms only see the constraint value c = (-1, -8, 20) and its Jacobian J (:269-287).  The same
# numbers therefore pin an affine device constraint with that J and e = c - J [x;u].  The device
# has no (3,2) model compiled in, so the knot is embedded in the linear (4,2) model with a fourth
# state that is identically zero (an extra zero column of J).
C_REF = np.array([-1.0, -8.0, 20.0])
J_REF = np.array([[-2.0, 0, 2, 0, 0, 0], [4, 4, 4, 0, 0, 0], [0, 0, 0, 0, 1, 1.0]])   # p x (4 + 2)
LXX_EQ = np.array([24.0, 19.2, 14.399999999999999, 19.2, 19.2, 19.2, 14.399999999999999, 19.2, 24.0]).reshape(3, 3)
HESS_SOC_OUT = np.array([
    13.121659323998685, 9.632047409257103, 6.142435494515529, 2.3820953755839365, 2.3820953755839365,
    9.632047409257108, 9.600915640264486, 9.569783871271873, 2.399740526514188, 2.399740526514188,
    6.142435494515531, 9.569783871271868, 12.997132248028219, 2.417385677444439, 2.417385677444439,
    2.382095375583937, 2.3997405265141882, 2.4173856774444396, 0.6, 0.6,
    2.382095375583937, 2.3997405265141882, 2.4173856774444396, 0.6, 0.6]).reshape(5, 5)
RHO = 1.2


def knot_on_gpu(cone, z):
    """KnotPointConstraintTest::InitializeKnotPoint (knotpoint_data_test.cpp:143-212) on the device;
    returns the AL parts: cost, lx, lu (constraint contribution only), full 5x5 constraint Hessian."""
    n, m, N, B = 4, 2, 1, 3
    x, u = np.array([2.0, 2, 2, 0]), np.array([10.0, 10])
    e = C_REF - J_REF @ np.concatenate([x, u])
    A = np.eye(n)
    Bm = np.zeros((n, m)); Bm[0, 0] = 1; Bm[1, 1] = 1; Bm[2, :] = 1.0
    s = BatchSolver(N, B)
    s.SetDimension(n, m)
    s.SetTimeStep(0.01)
    s.SetLinearDynamics(A, Bm)
    s.SetDiagonalCost(np.ones(n), np.ones(m), np.zeros(n), np.zeros(m), 0.0, 0, N + 1)
    s.SetConstraintAffine(cone, J_REF, e, 0, 1)
    s.SetInitialState(x)
    s.Initialize()
    s.SetState(x, 0, 1)
    s.SetState(np.zeros(n), 1, 2)
    s.SetInput(u, 0, 1)
    s.SetDualGeneric(0, 0, np.asarray(z, dtype=float))
    s.SetPenalty(RHO)
    s.KnotEval()
    assert np.allclose(s.GetField("constraint_val")[:, 0], C_REF, rtol=0, atol=1e-13)     # :269
    cost = s.CalcCost() - (0.5 * x @ x + 0.5 * u @ u)           # the AL term alone
    lx = s.GetField("lx")[:, 0] - x                               # original gradient Qx + q = x
    lu = s.GetField("lu")[:, 0] - u
    G = np.zeros((B, 5, 5))
    lxx = s.GetField("lxx")[:, 0].reshape(B, n, n).transpose(0, 2, 1) - np.eye(n)
    luu = s.GetField("luu")[:, 0].reshape(B, m, m).transpose(0, 2, 1) - np.eye(m)
    lux = s.GetField("lux")[:, 0].reshape(B, n, m).transpose(0, 2, 1)
    assert np.all(lxx[:, 3, :] == 0) and np.all(lxx[:, :, 3] == 0) and np.all(lux[:, :, 3] == 0)
    G[:, :3, :3], G[:, 3:, 3:], G[:, 3:, :3] = lxx[:, :3, :3], luu, lux[:, :, :3]
    G[:, :3, 3:] = lux[:, :, :3].transpose(0, 2, 1)
    z_est, z_proj = s.GetField("z_est")[:, 0], s.GetField("z_proj")[:, 0]
    s.close()
    for a in (cost, lx, lu, G, z_est, z_proj):   # every problem of the batch computes the same bits
        assert np.array_equal(a[0], a[1]) and np.array_equal(a[0], a[2])
    assert lx[0, 3] == 0
    return cost[0], lx[0, :3], lu[0], G[0], z_est[0], z_proj[0]


def test_knot_inequality_goldens_on_gpu():      # knotpoint_data_test.cpp:233-288
    z = np.array([-1, 4, 10.1])
    cost, lx, lu, G, z_est, z_proj = knot_on_gpu(INEQ, z)
    zt = np.minimum(z - RHO * C_REF, 0)
    assert np.allclose(z_est, z - RHO * C_REF, rtol=0, atol=1e-14) and np.allclose(z_proj, zt, rtol=0, atol=1e-14)
    assert cost == pytest.approx(zt @ zt / (2 * RHO), abs=1e-10)
    assert np.linalg.norm(lx) < 1e-10 and np.linalg.norm(lu - [13.9, 13.9]) < 1e-10
    assert np.all(G[:3, :] == 0) and np.allclose(G[3:, 3:], 1.2)


def test_knot_equality_goldens_on_gpu():        # :290-344
    z = np.array([-1, 4, 10.1])
    cost, lx, lu, G, z_est, z_proj = knot_on_gpu(EQ, z)
    zt = z - RHO * C_REF
    assert cost == pytest.approx(zt @ zt / (2 * RHO), abs=1e-10)
    assert np.linalg.norm(lx - [-54, -54.4, -54.8]) < 1e-10 and np.linalg.norm(lu - [13.9, 13.9]) < 1e-10
    assert np.linalg.norm(G[:3, :3] - LXX_EQ) < 1e-13 and np.all(G[3:, :3] == 0) and np.allclose(G[3:, 3:], 1.2)


def test_knot_soc_out_of_cone_goldens_on_gpu():  # :346-405
    cost, lx, lu, G, *_ = knot_on_gpu(SOC, [-1, 4, 30.0])
    assert cost == pytest.approx(80.04534293850527, abs=1e-10)
    assert np.linalg.norm(lx - [-38.910476877919685, -39.19870263257094, -39.4869283872222]) < 1e-10
    assert np.linalg.norm(lu - [-9.800735254367721, -9.800735254367721]) < 1e-10
    assert np.linalg.norm(G - HESS_SOC_OUT) < 1e-6


def test_knot_soc_below_cone_goldens_on_gpu():   # :407-462
    cost, lx, lu, G, *_ = knot_on_gpu(SOC, [-1, 4, 10.1])
    assert cost == pytest.approx(0.0, abs=1e-10)
    assert np.linalg.norm(lx) < 1e-10 and np.linalg.norm(lu) < 1e-10 and np.linalg.norm(G) < 1e-6


def test_knot_soc_in_cone_goldens_on_gpu():      # :464-524
    cost, lx, lu, G, *_ = knot_on_gpu(SOC, [-1, 4, 100.0])
    assert cost == pytest.approx(2483.75, abs=1e-10)
    assert np.linalg.norm(lx - [-54, -54.4, -54.8]) < 1e-10 and np.linalg.norm(lu - [-76, -76]) < 1e-10
    assert np.linalg.norm(G[:3, :3] - LXX_EQ) < 1e-13 and np.all(G[3:, :3] == 0) and np.allclose(G[3:, 3:], 1.2)


def test_bound_setters_reproduce_control_box(oracle):
    """SetInputUpperBound / SetInputLowerBound semantics (INEQUALITY rows u - u_max <= 0,
    u_min - u <= 0) through the C ABI: the double-integrator control-box problem converges in the
    reference's 5 iterations with u0 = -1 (double_integrator_test.cpp:367-374)."""
    P = PR.double_integrator(N=10, variant="ubox")
    n, m = P.n, P.m
    s = BatchSolver(P.N, 1)
    s.SetDimension(n, m)
    s.SetTimeStep(P.h)
    s.SetExplicitDynamics(P.model_id, P.model_params)
    altro_b200.solver.set_cost(s, P)
    goal = [c for c in P.constraints if c.cone == EQ][0]
    s.SetConstraint(goal.cone, goal.idx, goal.scale, goal.off, goal.k_start, goal.k_stop)
    # upper bound rows u_i - 1 <= 0, then lower bound rows -1 - u_i <= 0, as two constraints
    s.SetConstraint(INEQ, [n, n + 1], [1.0, 1.0], [-1.0, -1.0], 0, P.N)
    s.SetConstraint(INEQ, [n, n + 1], [-1.0, -1.0], [-1.0, -1.0], 0, P.N)
    s.SetInitialState(P.x0)
    s.Initialize()
    s.SetInput(P.U0)
    s.SetOptions(default_options(**P.options))
    st = s.Solve()
    assert st[0] == 0 and s.GetIterations()[0] == 5
    u0 = s.GetInputs()[0, 0]
    assert np.abs(u0 + 1.0).max() < 1e-4
    assert np.linalg.norm(s.GetStates()[0, -1]) < 1e-4
    s.close()


def test_per_knot_time_steps_match_oracle(oracle):
    """SetTimeStep(h, k_start, k_stop) (altro_solver.cpp:49-63) with two different steps over the
    horizon on the nonlinear pendulum (Jacobians, rollouts, packed [A B] all depend on h_k): the
    pipeline against the oracle, and bit for bit against the persistent twin."""
    rng = np.random.default_rng(21)
    n, m, N, B = 2, 1, 40, 40
    h1, h2, ksw = np.float32(0.05), np.float32(0.075), 17   # (0.1 makes this search fail on CPU and GPU alike)
    Qd, Rd, Qf = np.full(n, 1e-2), np.full(m, 1e-3), np.full(n, 1.0)
    xf = np.array([np.pi, 0.0])
    x0 = rng.uniform(-0.5, 0.5, size=(B, n))

    def build(mode):
        s = BatchSolver(N, B)
        s.SetDimension(n, m)
        s.SetTimeStep(float(h1), 0, ksw)
        s.SetTimeStep(float(h2), ksw, N)
        s.SetExplicitDynamics(PR.MODEL_PENDULUM, [])
        s.SetInitialState(x0)
        s.SetLQRCost(Qd, Rd, xf, np.zeros(m), 0, N)
        s.SetLQRCost(Qf, Rd, xf, np.zeros(m), N, N + 1)
        s.Initialize()
        s.SetInput(np.array([0.1]))
        s.SetOptions(default_options(iterations_max=60))
        s.SetSolveMode(mode)
        s.Solve()
        return s

    s = build(0)
    twin = build(1)
    assert np.array_equal(s.GetStates(), twin.GetStates()) and np.array_equal(s.GetInputs(), twin.GetInputs())
    assert np.array_equal(s.GetIterations(), twin.GetIterations())
    twin.close()
    for b in range(0, B, 5):
        o = oracle.OracleSolver(N, n, m)
        o.SetTimeStep(float(h1), 0, ksw)
        o.SetTimeStep(float(h2), ksw, N)
        o.SetModel(oracle.MODEL_PENDULUM, [])
        o.SetInitialState(x0[b])
        for k in range(N):
            o.SetLQRCost(k, Qd, Rd, xf, np.zeros(m))
        o.SetLQRCost(N, Qf, Rd, xf, np.zeros(m))
        o.Initialize()
        o.SetInput(np.array([0.1]))
        o.SetOptions(oracle.default_options(iterations_max=60))
        st = o.Solve()
        assert st == 0 and o.GetIterations() > 4
        assert_same_solve(s, o, b, st)
    # a horizon with a knot that never got a step is refused like the reference refuses it
    s2 = BatchSolver(N, 2)
    s2.SetDimension(n, m)
    s2.SetTimeStep(0.05, 0, N - 1)
    s2.SetExplicitDynamics(PR.MODEL_PENDULUM, [])
    s2.SetInitialState(x0[:2])
    s2.SetLQRCost(Qd, Rd, xf, np.zeros(m), 0, N + 1)
    with pytest.raises(altro_b200.AltroB200Error):
        s2.Initialize()
    s2.close()
    s.close()
