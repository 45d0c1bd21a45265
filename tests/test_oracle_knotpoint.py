"""Oracle AL terms vs src/altro/solver/test/knotpoint_data_test.cpp:136-524 goldens."""
import numpy as np
import pytest

EQ, ID, INEQ, SOC = 0, 1, 2, 3


def make_knot(O, cone, z):
    """KnotPointConstraintTest::InitializeKnotPoint (knotpoint_data_test.cpp:143-212)."""
    n, m, p = 3, 2, 3
    c1, c2 = np.array([1.0, 2, 3]), np.array([4.0, 4, 4])
    r1, r2 = 1.0, 2.0

    def con(x, u):
        return [r1 * r1 - np.sum((x - c1) ** 2), r2 * r2 - np.sum((x - c2) ** 2), u[0] + u[1]]

    def jac(x, u):
        J = np.zeros((p, n + m))
        J[0, :3] = -2 * (x - c1)
        J[1, :3] = -2 * (x - c2)
        J[2, 3:] = 1.0
        return J

    s = O.OracleSolver(1, n, m)
    s.SetTimeStep(0.01)
    A = np.eye(n)
    B = np.zeros((n, m)); B[0, 0] = 1; B[1, 1] = 1; B[2, :] = 1.0
    s.SetLinearDynamics(0, A, B, None)
    s.SetDiagonalCost(0, np.ones(n), np.ones(m), np.zeros(n), np.zeros(m), 0.0)
    s.SetDiagonalCost(1, np.ones(n), np.ones(m), np.zeros(n), np.zeros(m), 0.0)
    s.AddCallbackConstraint(0, cone, p, con, jac)
    assert s.Initialize() == 0
    s.set(0, "x_", [2.0, 2, 2])
    s.set(0, "u_", [10.0, 10])
    s.SetPenalty(1.2)
    s.SetDual(0, 0, z)
    return s, n, m, p


C_EXPECTED = np.array([-1.0, -8.0, 20.0])
J_EXPECTED = np.array([[-2.0, 0, 2, 0, 0], [4, 4, 4, 0, 0], [0, 0, 0, 1, 1.0]])


def common_checks(s, n, m, p):
    s.KnotOp(0, "CalcConstraints")
    assert np.linalg.norm(s.get(0, "constraint_val_0") - C_EXPECTED) < 1e-6
    alcost = s.KnotCalcConstraintCosts(0)
    s.KnotOp(0, "CalcConstraintJacobians")
    assert np.linalg.norm(s.get(0, "constraint_jac_0", (p, n + m)) - J_EXPECTED) < 1e-6
    return alcost


def grads_hess(s, n, m):
    s.set(0, "lx_", np.zeros(n)); s.set(0, "lu_", np.zeros(m))
    s.KnotOp(0, "CalcConstraintCostGradients")
    lx, lu = s.get(0, "lx_"), s.get(0, "lu_")
    s.set(0, "lxx_", np.zeros((n, n))); s.set(0, "luu_", np.zeros((m, m))); s.set(0, "lux_", np.zeros((m, n)))
    s.KnotOp(0, "CalcConstraintCostHessians")
    return lx, lu, s.get(0, "lxx_", (n, n)), s.get(0, "luu_", (m, m)), s.get(0, "lux_", (m, n))


def test_inequality(oracle):
    z = np.array([-1, 4, 10.1]); rho = 1.2
    s, n, m, p = make_knot(oracle, INEQ, z)
    alcost = common_checks(s, n, m, p)
    zt = np.minimum(z - rho * C_EXPECTED, 0)
    assert alcost == pytest.approx(zt @ zt / (2 * rho), abs=1e-10)
    lx, lu, lxx, luu, lux = grads_hess(s, n, m)
    assert np.linalg.norm(lx) < 1e-10 and np.linalg.norm(lu - [13.9, 13.9]) < 1e-10
    assert np.all(s.get(0, "proj_hess_0") == 0) and np.all(lxx == 0) and np.all(lux == 0)
    assert np.allclose(luu, 1.2)


LXX_EQ = np.array([24.0, 19.2, 14.399999999999999, 19.2, 19.2, 19.2, 14.399999999999999, 19.2, 24.0]).reshape(3, 3)


def test_equality(oracle):
    z = np.array([-1, 4, 10.1]); rho = 1.2
    s, n, m, p = make_knot(oracle, EQ, z)
    alcost = common_checks(s, n, m, p)
    zt = z - rho * C_EXPECTED
    assert alcost == pytest.approx(zt @ zt / (2 * rho), abs=1e-10)
    lx, lu, lxx, luu, lux = grads_hess(s, n, m)
    assert np.linalg.norm(lx - [-54, -54.4, -54.8]) < 1e-10
    assert np.linalg.norm(lu - [13.9, 13.9]) < 1e-10
    assert np.linalg.norm(lxx - LXX_EQ) < 1e-13 and np.all(lux == 0) and np.allclose(luu, 1.2)


def test_soc_out_of_cone(oracle):
    s, n, m, p = make_knot(oracle, SOC, np.array([-1, 4, 30.0]))
    alcost = common_checks(s, n, m, p)
    assert alcost == pytest.approx(80.04534293850527, abs=1e-10)
    lx, lu, *_ = grads_hess(s, n, m)
    assert np.linalg.norm(lx - [-38.910476877919685, -39.19870263257094, -39.4869283872222]) < 1e-10
    assert np.linalg.norm(lu - [-9.800735254367721, -9.800735254367721]) < 1e-10
    hess_expected = np.array([
        13.121659323998685, 9.632047409257103, 6.142435494515529, 2.3820953755839365, 2.3820953755839365,
        9.632047409257108, 9.600915640264486, 9.569783871271873, 2.399740526514188, 2.399740526514188,
        6.142435494515531, 9.569783871271868, 12.997132248028219, 2.417385677444439, 2.417385677444439,
        2.382095375583937, 2.3997405265141882, 2.4173856774444396, 0.6, 0.6,
        2.382095375583937, 2.3997405265141882, 2.4173856774444396, 0.6, 0.6]).reshape(5, 5)
    assert np.linalg.norm(s.get(0, "constraint_hess_0", (n + m, n + m)) - hess_expected) < 1e-6


def test_soc_below_cone(oracle):
    s, n, m, p = make_knot(oracle, SOC, np.array([-1, 4, 10.1]))
    alcost = common_checks(s, n, m, p)
    assert alcost == pytest.approx(0.0, abs=1e-10)
    lx, lu, *_ = grads_hess(s, n, m)
    assert np.linalg.norm(lx) < 1e-10 and np.linalg.norm(lu) < 1e-10
    assert np.linalg.norm(s.get(0, "constraint_hess_0")) < 1e-6


def test_soc_in_cone(oracle):
    s, n, m, p = make_knot(oracle, SOC, np.array([-1, 4, 100.0]))
    alcost = common_checks(s, n, m, p)
    assert alcost == pytest.approx(2483.75, abs=1e-10)
    lx, lu, lxx, luu, lux = grads_hess(s, n, m)
    assert np.linalg.norm(lx - [-54, -54.4, -54.8]) < 1e-10
    assert np.linalg.norm(lu - [-76, -76]) < 1e-10
    assert np.all(s.get(0, "proj_hess_0") == 0)
    assert np.linalg.norm(lxx - LXX_EQ) < 1e-13 and np.all(lux == 0) and np.allclose(luu, 1.2)


def test_cost_expansion_and_init_errors(oracle):
    """knotpoint_data_test.cpp:43-134 (values; the error sequence lives in the C++ facade)."""
    O = oracle
    n, m = 4, 2
    x = np.array([0.1, 0.2, -0.3, -1.1]); u = np.array([10.1, -20.2])
    Qd, Rd = np.full(n, 1.1), np.full(m, 0.1)
    q, r = np.full(n, 0.01), np.full(m, 0.001)
    s = O.OracleSolver(1, n, m)
    assert s.Initialize() != 0           # nothing set yet
    s.SetTimeStep(0.01)
    A = np.eye(n); A[0, 2] = A[1, 3] = 0.01
    B = np.zeros((n, m)); B[2, 0] = B[3, 1] = 0.01
    s.SetLinearDynamics(0, A, B, np.zeros(n))
    s.SetDiagonalCost(0, Qd, Rd, q, r, 10.5)
    s.SetDiagonalCost(1, Qd, None, q, None, 10.5)
    assert s.Initialize() == 0
    assert np.allclose(np.diag(s.get(0, "lxx_", (n, n))), Qd)
    assert np.allclose(np.diag(s.get(0, "luu_", (m, m))), Rd)
    s.set(0, "x_", x); s.set(0, "u_", u)
    s.KnotOp(0, "CalcCostGradient"); s.KnotOp(0, "CalcCostHessian")
    assert np.allclose(s.get(0, "lx_"), Qd * x + q) and np.allclose(s.get(0, "lu_"), Rd * u + r)
    assert np.all(s.get(0, "lux_") == 0)
    s.set(1, "x_", x)
    s.KnotOp(1, "CalcCostGradient"); s.KnotOp(1, "CalcCostHessian")
    assert np.allclose(s.get(1, "lxx_", (n, n)), np.diag(Qd)) and np.allclose(s.get(1, "lx_"), Qd * x + q)
