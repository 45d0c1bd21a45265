"""Oracle cones + tvlqr vs the reference goldens:
src/altro/solver/test/cones_test.cpp, src/tvlqr/test/tvlqr_test.cpp."""
import ctypes as C

import numpy as np

EQ, ID, INEQ, SOC = 0, 1, 2, 3


def test_projections(oracle):
    O = oracle
    x = np.array([0.1, -0.5, 0.2, 0.0])
    assert np.allclose(O.conic_projection(EQ, x), 0, atol=1e-10)           # cones_test.cpp:9-18
    assert np.allclose(O.conic_projection(INEQ, x), [0, -0.5, 0, 0], atol=1e-10)  # :20-29
    assert np.allclose(O.conic_projection(ID, x), x, atol=1e-10)           # :31-40
    mag = np.linalg.norm(x)                                                # :42-69
    xs = x.copy(); xs[3] = mag * 1.1
    assert np.allclose(O.conic_projection(SOC, xs), xs, atol=1e-10)
    xs[3] = -mag * 1.1
    assert np.allclose(O.conic_projection(SOC, xs), 0, atol=1e-10)
    xs[3] = mag * 0.9
    assert np.linalg.norm(O.conic_projection(SOC, xs) - [0.095, -0.475, 0.19, 0.5203364296299079]) < 1e-10


def test_projection_jacobians(oracle):
    O = oracle
    x = np.array([0.1, -0.5, 0.2, 0.0])
    assert np.allclose(O.conic_projection_jacobian(EQ, x), 0)
    assert np.allclose(O.conic_projection_jacobian(ID, x), np.eye(4))
    assert np.allclose(O.conic_projection_jacobian(INEQ, x), np.diag([0, 1, 0, 1.0]))  # :95-104
    mag = np.linalg.norm(x)
    xs = x.copy(); xs[3] = mag * 1.1
    assert np.allclose(O.conic_projection_jacobian(SOC, xs), np.eye(4))
    xs[3] = -mag * 1.1
    assert np.allclose(O.conic_projection_jacobian(SOC, xs), 0)
    xs[3] = mag * 0.9                                                      # :128-135
    exp = np.array([0.9349999999999999, 0.07499999999999998, -0.029999999999999995, 0.09128709291752768,
                    0.07499999999999998, 0.5750000000000001, 0.14999999999999997, -0.4564354645876384,
                    -0.029999999999999995, 0.14999999999999997, 0.89, 0.18257418583505536,
                    0.09128709291752768, -0.45643546458763834, 0.18257418583505536, 0.5]).reshape(4, 4)
    assert np.linalg.norm(O.conic_projection_jacobian(SOC, xs) - exp) < 1e-10


def test_projection_hessians(oracle):
    O = oracle
    x = np.array([0.1, -0.5, 0.2, 0.0])
    b = np.array([10, 20, -30, 40.0])
    for cone in (EQ, ID, INEQ):
        assert np.allclose(O.conic_projection_hessian(cone, x, b), 0)
    mag = np.linalg.norm(x)
    xs = x.copy(); xs[3] = mag * 1.1
    assert np.allclose(O.conic_projection_hessian(SOC, xs, b), 0)
    xs[3] = -mag * 1.1
    assert np.allclose(O.conic_projection_hessian(SOC, xs, b), 0)
    xs[3] = mag * 0.9                                                      # :201-209
    exp = np.array([52.54767592811069, 21.83580619450183, -5.434322477800736, 13.69306393762915,
                    21.83580619450183, 2.3358061945018775, 6.1716123890036805, -4.564354645876377,
                    -5.434322477800736, 6.1716123890036805, 63.146192211409584, -18.257418583505533,
                    13.69306393762915, -4.564354645876377, -18.257418583505533, 0.0]).reshape(4, 4)
    H = O.conic_projection_hessian(SOC, xs, b)
    assert np.linalg.norm(H - exp) < 1e-10
    assert np.linalg.norm(H - H.T) < 1e-10


def tvlqr_problem(h):
    """The double-integrator LQR problem of tvlqr_test.cpp:15-66 (also solver_impl_test.cpp:19-56)."""
    n, m, N, dim = 4, 2, 10, 2
    hf = np.float32(h)
    b = float(np.float32(hf * hf / np.float32(2)))
    hd = float(hf)
    A = np.eye(n); A[0, 2] = hd; A[1, 3] = hd
    B = np.zeros((n, m)); B[0, 0] = b; B[1, 1] = b; B[2, 0] = hd; B[3, 1] = hd
    xeq = np.array([1.0, 2, 0, 0]); ueq = np.zeros(2)
    f = A @ xeq + B @ ueq      # discrete_double_integrator_dynamics(f, xeq, ueq)
    return dict(n=n, m=m, N=N, A=A, B=B, f=f, Qd=np.full(n, 1.1), Rd=np.full(m, 0.1),
                Qfd=np.full(n, 110.0), q=np.full(n, 0.01), r=np.full(m, 0.001),
                x0=np.array([10.5, -20.5, -4, 5.0]))


K0_EXPECTED = np.array([[0.7753129718046554, 0.0, 5.840445640045901, 0.0],
                        [0.0, 0.7753129718046554, 0.0, 5.840445640045901]])
D0_EXPECTED = np.array([-7.634078625343007, -15.256221385516275])
XN_EXPECTED = np.array([20.165445369740308, -0.13732391651279308, -2.3724421496097037, 2.3113121303468707])
YN_EXPECTED = np.array([2218.2089906714345, -15.09563081640724, -260.9586364570674, 254.2543343381558])


def run_tvlqr(O, pr, is_diag=True):
    L = O.lib()
    n, m, N = pr["n"], pr["m"], pr["N"]
    dp = O.dptr

    def table(arrs):
        keep = [np.ascontiguousarray(a, dtype=float) for a in arrs]
        t = (dp * len(keep))(*[a.ctypes.data_as(dp) for a in keep])
        return t, keep

    colmajor = lambda M: np.asarray(M).T.reshape(-1).copy()
    A, kA = table([colmajor(pr["A"])] * N)
    Bt, kB = table([colmajor(pr["B"])] * N)
    f, kf = table([pr["f"]] * N)
    if is_diag:
        Q, kQ = table([pr["Qd"]] * N + [pr["Qfd"]])
        R, kR = table([pr["Rd"]] * N)
    else:
        Q, kQ = table([colmajor(np.diag(pr["Qd"]))] * N + [colmajor(np.diag(pr["Qfd"]))])
        R, kR = table([colmajor(np.diag(pr["Rd"]))] * N)
    H, kH = table([np.zeros(m * n)] * N)
    q, kq = table([pr["q"]] * (N + 1))
    r, kr = table([pr["r"]] * N)
    outs = {}
    for name, size, cnt in [("K", m * n, N), ("d", m, N), ("P", n * n, N + 1), ("p", n, N + 1),
                            ("Qxx", n * n, N), ("Quu", m * m, N), ("Qux", m * n, N), ("Qx", n, N),
                            ("Qu", m, N), ("Qxx_tmp", n * n, N), ("Quu_tmp", m * m, N),
                            ("Qux_tmp", m * n, N), ("Qx_tmp", n, N), ("Qu_tmp", m, N),
                            ("x", n, N + 1), ("u", m, N), ("y", n, N + 1)]:
        outs[name] = table([np.zeros(size) for _ in range(cnt)])
    nx = (C.c_int * (N + 1))(*([n] * (N + 1)))
    nu = (C.c_int * N)(*([m] * N))
    dV = np.zeros(2)
    T = lambda k: outs[k][0]
    res = L.oracle_tvlqr_backward_pass(nx, nu, N, A, Bt, f, Q, R, H, q, r, C.c_double(0.0), T("K"), T("d"),
                                       T("P"), T("p"), dV.ctypes.data_as(dp), T("Qxx"), T("Quu"),
                                       T("Qux"), T("Qx"), T("Qu"), T("Qxx_tmp"), T("Quu_tmp"),
                                       T("Qux_tmp"), T("Qx_tmp"), T("Qu_tmp"), 0, int(is_diag))
    x0 = np.ascontiguousarray(pr["x0"])
    res2 = L.oracle_tvlqr_forward_pass(nx, nu, N, A, Bt, f, T("K"), T("d"), T("P"), T("p"),
                                       x0.ctypes.data_as(dp), T("x"), T("u"), T("y"))
    K0 = outs["K"][1][0].reshape(n, m).T
    return res, res2, K0, outs["d"][1][0], outs["x"][1][N], outs["y"][1][N], dV, outs


def test_tvlqr_goldens_float_h(oracle):
    """tvlqr_test.cpp:185-213 with the reference's float h: K0,d0 to 1e-6, xN 1e-6, yN 1e-5."""
    res, res2, K0, d0, xN, yN, dV, _ = run_tvlqr(oracle, tvlqr_problem(np.float32(0.01)))
    assert res == -1 and res2 == -1          # TVLQR_SUCCESS, tvlqr.h:11
    assert np.linalg.norm(K0 - K0_EXPECTED) < 1e-6
    assert np.linalg.norm(d0 - D0_EXPECTED) < 1e-6
    assert np.abs(xN - XN_EXPECTED).max() < 1e-6
    assert np.abs(yN - YN_EXPECTED).max() < 1e-5


def test_tvlqr_goldens_double_h_exact(oracle):
    """With a double h the goldens are reproduced to ~machine precision."""
    pr = tvlqr_problem(0.01)
    h = 0.01
    A = np.eye(4); A[0, 2] = h; A[1, 3] = h
    B = np.zeros((4, 2)); B[0, 0] = h * h / 2; B[1, 1] = h * h / 2; B[2, 0] = h; B[3, 1] = h
    pr.update(A=A, B=B, f=A @ np.array([1.0, 2, 0, 0]))
    res, _, K0, d0, xN, yN, dV, _ = run_tvlqr(oracle, pr)
    assert np.abs(K0 - K0_EXPECTED).max() < 1e-12
    assert np.abs(d0 - D0_EXPECTED).max() < 1e-11
    assert np.abs(xN - XN_EXPECTED).max() < 1e-11
    assert np.abs(yN - YN_EXPECTED).max() < 1e-9
    assert np.allclose(dV, [-80.823, 40.412], atol=2e-3)


def test_tvlqr_dense_equals_diag(oracle):
    pr = tvlqr_problem(np.float32(0.01))
    a = run_tvlqr(oracle, pr, True)
    b = run_tvlqr(oracle, pr, False)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[4])


def test_tvlqr_mem_size(oracle):
    """tvlqr_test.cpp:67-72,167: size equals the hand-laid layout."""
    n, m, N = 4, 2, 10
    nx = (C.c_int * (N + 1))(*([n] * (N + 1)))
    nu = (C.c_int * N)(*([m] * N))
    per_k = (n + m + n) + (n * n + n * m + n) + (n + n + m + m) + (m * n + m) + (n * n + n) \
        + 2 * (n * n + m * m + m * n + n + m)
    term = n + n + n + n + n * n + n + 2
    assert oracle.lib().oracle_tvlqr_total_mem_size(nx, nu, N, 1) == (N * per_k + term) * 8
    assert oracle.lib().oracle_tvlqr_total_mem_size(None, nu, N, 1) == 0


def test_tvlqr_cholesky_failure_returns_knot(oracle):
    """tvlqr.cpp:162-164: returns the failing knot index."""
    pr = tvlqr_problem(np.float32(0.01))
    pr["Rd"] = np.full(2, -1e6)
    res = run_tvlqr(oracle, pr)[0]
    assert res == pr["N"] - 1
