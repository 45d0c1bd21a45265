"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (altro_b200/) never imports this.
See oracle/altro_oracle.h for the reference citations of every entry point.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

EQUALITY, IDENTITY, INEQUALITY, SOC = 0, 1, 2, 3
STATUS_SUCCESS, STATUS_UNSOLVED, STATUS_MAX_ITERATIONS = 0, 1, 2
MODEL_LINEAR, MODEL_DI, MODEL_PENDULUM, MODEL_BICYCLE4, MODEL_BICYCLE5, MODEL_CHAIN = range(6)
MAX_CON, MAX_CON_DIM = 8, 16

OPS = {name: i for i, name in enumerate([
    "CalcConstraints", "CalcConstraintJacobians", "CalcProjectedDuals", "CalcConicJacobians",
    "CalcConicHessians", "CalcCostGradient", "CalcCostHessian", "CalcDynamicsExpansion",
    "CalcConstraintCostGradients", "CalcConstraintCostHessians", "CalcOriginalCostGradient",
    "CalcOriginalCostHessian"])}

dptr = C.POINTER(C.c_double)
iptr = C.POINTER(C.c_int)


class Options(C.Structure):
    _fields_ = [("iterations_max", C.c_int), ("tol_primal_feasibility", C.c_double),
                ("tol_stationarity", C.c_double), ("tol_meritfun_gradient", C.c_double),
                ("penalty_initial", C.c_double), ("penalty_scaling", C.c_double),
                ("penalty_max", C.c_double), ("use_backtracking_linesearch", C.c_int),
                ("ls_c1", C.c_double), ("ls_c2", C.c_double)]


class ConSpec(C.Structure):
    _fields_ = [("k_start", C.c_int), ("k_stop", C.c_int), ("cone", C.c_int), ("dim", C.c_int),
                ("idx", C.c_int * MAX_CON_DIM), ("scale", C.c_double * MAX_CON_DIM),
                ("off", C.c_double * MAX_CON_DIM), ("off_b", dptr)]


class BatchSpec(C.Structure):
    _fields_ = [("N", C.c_int), ("n", C.c_int), ("m", C.c_int), ("B", C.c_int), ("h", C.c_float),
                ("model_id", C.c_int), ("model_params", C.c_double * 8),
                ("Qd", dptr), ("Rd", dptr), ("ref_mode", C.c_int),
                ("q", dptr), ("r", dptr), ("c", dptr), ("xref", dptr), ("uref", dptr),
                ("offsets", iptr), ("T", C.c_int), ("x0", dptr), ("U0", dptr),
                ("U0_per_problem", C.c_int), ("ncon", C.c_int), ("con", ConSpec * MAX_CON),
                ("opts", Options)]


class BatchResult(C.Structure):
    _fields_ = [("X", dptr), ("U", dptr), ("Y", dptr), ("status", iptr), ("iters", iptr),
                ("merit_evals", C.POINTER(C.c_long)), ("cost", dptr), ("stat", dptr),
                ("feas", dptr)]


MERIT_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_double, dptr, dptr)
CON_CB = C.CFUNCTYPE(None, C.c_void_p, dptr, dptr, dptr)


class LineSearch(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("alpha_max", C.c_double), ("beta_increase", C.c_double),
                ("beta_decrease", C.c_double), ("min_interval_size", C.c_double),
                ("try_cubic_first", C.c_int), ("use_backtracking_linesearch", C.c_int),
                ("c1", C.c_double), ("c2", C.c_double), ("return_code", C.c_int),
                ("n_iters", C.c_int), ("phi0", C.c_double), ("phi", C.c_double),
                ("phi_lo", C.c_double), ("phi_hi", C.c_double), ("dphi0", C.c_double),
                ("dphi", C.c_double), ("dphi_lo", C.c_double), ("dphi_hi", C.c_double),
                ("sufficient_decrease", C.c_int), ("curvature", C.c_int)]


class Spline(C.Structure):
    _fields_ = [("x0", C.c_double), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("d", C.c_double)]


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in
            ("altro_oracle.c", "linesearch_port.c", "models.c", "batch.c", "altro_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so", "liboracle_fma.so"], check=True,
                       capture_output=True)
    if os.path.isdir("/root/reference/src/linesearch"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return so


def use_variant(name="liboracle.so"):
    """Switch the loaded oracle build (liboracle.so | liboracle_fma.so); test use only."""
    global _LIB, _SO_NAME
    _LIB = None
    _SO_NAME = name


_SO_NAME = "liboracle.so"


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, _SO_NAME)
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int] * 3
        L.oracle_batch_make_solver.restype = C.c_void_p
        L.oracle_batch_solve.restype = C.c_double
        L.oracle_batch_solve.argtypes = [C.POINTER(BatchSpec), C.c_int, C.c_int, C.c_int,
                                         C.POINTER(BatchResult)]
        for f in ("oracle_calc_cost", "oracle_stationarity", "oracle_feasibility",
                  "oracle_knot_calc_cost", "oracle_knot_calc_constraint_costs",
                  "oracle_knot_calc_violations", "oracle_get_final_phi", "oracle_ls_run",
                  "oracle_spline_argmin"):
            getattr(L, f).restype = C.c_double
        L.oracle_get_merit_evals.restype = C.c_long
        L.oracle_set_time_step.argtypes = [C.c_void_p, C.c_float]
        L.oracle_set_time_step_range.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        L.oracle_set_penalty.argtypes = [C.c_void_p, C.c_double]
        L.oracle_merit_function.argtypes = [C.c_void_p, C.c_double, dptr, dptr]
        L.oracle_ls_run.argtypes = [C.POINTER(LineSearch), MERIT_CB, C.c_void_p, C.c_double,
                                    C.c_double, C.c_double]
        L.oracle_spline_from2points.argtypes = [C.POINTER(Spline)] + [C.c_double] * 6
        L.oracle_spline_argmin.argtypes = [C.POINTER(Spline), iptr]
        L.oracle_model_dynamics.argtypes = [C.c_int, dptr, dptr, dptr, dptr, C.c_float]
        L.oracle_model_jacobian.argtypes = [C.c_int, dptr, dptr, dptr, dptr, C.c_float]
        L.oracle_model_continuous.argtypes = [C.c_int, dptr, dptr, dptr, dptr]
        L.oracle_model_continuous_jacobian.argtypes = [C.c_int, dptr, dptr, dptr, dptr]
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own compiled line search (oracle/_ref), or None if not built."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "liblinesearch_ref.so")
        if not os.path.exists(so):
            return None
        R = C.CDLL(so)
        R.ref_linesearch_run.restype = C.c_double
        R.ref_linesearch_run.argtypes = [MERIT_CB, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_int, C.c_int, iptr, iptr,
                                         dptr, dptr, iptr, iptr]
        R.ref_cubic_argmin.restype = C.c_double
        R.ref_cubic_argmin.argtypes = [C.c_double] * 6 + [iptr, iptr, dptr]
        _REF = R
    return _REF


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dptr)


def default_options(**kw):
    o = Options()
    lib().oracle_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def model_dynamics(model_id, params, x, u, h):
    x, xp = _d(x)
    u, up = _d(u)
    prm, pp = _d(list(params) + [0.0] * (8 - len(params)))
    xn = np.zeros_like(x)
    lib().oracle_model_dynamics(model_id, pp, xn.ctypes.data_as(dptr), xp, up, C.c_float(h))
    return xn


def model_jacobian(model_id, params, x, u, h):
    x, xp = _d(x)
    u, up = _d(u)
    prm, pp = _d(list(params) + [0.0] * (8 - len(params)))
    n, m = len(x), len(u)
    J = np.zeros(n * (n + m))
    lib().oracle_model_jacobian(model_id, pp, J.ctypes.data_as(dptr), xp, up, C.c_float(h))
    return J.reshape(n + m, n).T.copy()


def model_continuous(model_id, params, x, u):
    x, xp = _d(x)
    u, up = _d(u)
    prm, pp = _d(list(params) + [0.0] * (8 - len(params)))
    xd = np.zeros_like(x)
    lib().oracle_model_continuous(model_id, pp, xd.ctypes.data_as(dptr), xp, up)
    return xd


def model_continuous_jacobian(model_id, params, x, u):
    x, xp = _d(x)
    u, up = _d(u)
    prm, pp = _d(list(params) + [0.0] * (8 - len(params)))
    n, m = len(x), len(u)
    J = np.zeros(n * (n + m))
    lib().oracle_model_continuous_jacobian(model_id, pp, J.ctypes.data_as(dptr), xp, up)
    return J.reshape(n + m, n).T.copy()


def conic_projection(cone, x):
    x, xp = _d(x)
    px = np.zeros_like(x)
    lib().oracle_conic_projection(cone, len(x), xp, px.ctypes.data_as(dptr))
    return px


def conic_projection_jacobian(cone, x):
    x, xp = _d(x)
    p = len(x)
    J = np.zeros(p * p)
    lib().oracle_conic_projection_jacobian(cone, p, xp, J.ctypes.data_as(dptr))
    return J.reshape(p, p).T.copy()


def conic_projection_hessian(cone, x, b):
    x, xp = _d(x)
    b, bp = _d(b)
    p = len(x)
    H = np.zeros(p * p)
    lib().oracle_conic_projection_hessian(cone, p, xp, bp, H.ctypes.data_as(dptr))
    return H.reshape(p, p).T.copy()


def cubic_argmin(x1, y1, d1, x2, y2, d2):
    """(argmin, build_err, argmin_err, coeffs) via the oracle port."""
    p = Spline()
    e1 = lib().oracle_spline_from2points(C.byref(p), x1, y1, d1, x2, y2, d2)
    coeffs = (p.x0, p.a, p.b, p.c, p.d)
    if e1 != 0:
        return float("nan"), e1, -1, coeffs
    e2 = C.c_int(0)
    x = lib().oracle_spline_argmin(C.byref(p), C.byref(e2))
    return x, e1, e2.value, coeffs


def linesearch_run(merit, alpha0, phi0, dphi0, c1=1e-4, c2=0.9, try_cubic_first=False,
                   use_backtracking=False):
    """merit(alpha, want_derivative) -> (phi, dphi).  Returns dict of results (oracle port)."""
    ls = LineSearch()
    lib().oracle_ls_init(C.byref(ls))
    ls.c1, ls.c2 = c1, c2
    ls.try_cubic_first = int(try_cubic_first)
    ls.use_backtracking_linesearch = int(use_backtracking)
    alphas = []

    def cb(_ctx, alpha, phi_p, dphi_p):
        want = bool(dphi_p)
        alphas.append(alpha)
        phi, dphi = merit(alpha, want)
        phi_p[0] = phi
        if want:
            dphi_p[0] = dphi

    alpha = lib().oracle_ls_run(C.byref(ls), MERIT_CB(cb), None, alpha0, phi0, dphi0)
    return dict(alpha=alpha, status=ls.return_code, iters=ls.n_iters, phi=ls.phi, dphi=ls.dphi,
                sufficient_decrease=bool(ls.sufficient_decrease), curvature=bool(ls.curvature),
                alphas=alphas)


def ref_linesearch_run(merit, alpha0, phi0, dphi0, c1=1e-4, c2=0.9, try_cubic_first=False,
                       use_backtracking=False):
    """Same as linesearch_run but through the REFERENCE's compiled linesearch.cpp."""
    R = ref_lib()
    alphas = []

    def cb(_ctx, alpha, phi_p, dphi_p):
        want = bool(dphi_p)
        alphas.append(alpha)
        phi, dphi = merit(alpha, want)
        phi_p[0] = phi
        if want:
            dphi_p[0] = dphi

    st, it, sd, cv = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    ph, dph = C.c_double(), C.c_double()
    alpha = R.ref_linesearch_run(MERIT_CB(cb), None, alpha0, phi0, dphi0, c1, c2,
                                 int(try_cubic_first), int(use_backtracking), C.byref(st),
                                 C.byref(it), C.byref(ph), C.byref(dph), C.byref(sd), C.byref(cv))
    return dict(alpha=alpha, status=st.value, iters=it.value, phi=ph.value, dphi=dph.value,
                sufficient_decrease=bool(sd.value), curvature=bool(cv.value), alphas=alphas)


class OracleSolver:
    """Single-problem oracle solver; method names follow altro::ALTROSolver / SolverImpl."""

    def __init__(self, N, n, m):
        self.L = lib()
        self.N, self.n, self.m = N, n, m
        self.h = C.c_void_p(self.L.oracle_create(N, n, m))
        self._keep = []

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    # --- problem definition
    def SetOptions(self, opts):
        self.L.oracle_set_options(self.h, C.byref(opts))

    def SetTimeStep(self, h, k_start=None, k_stop=None):
        if k_start is None:
            self.L.oracle_set_time_step(self.h, C.c_float(h))
        else:
            self.L.oracle_set_time_step_range(self.h, C.c_float(h), k_start, k_stop)

    def SetModel(self, model_id, params=()):
        prm, pp = _d(list(params) + [0.0] * (8 - len(params)))
        self.L.oracle_set_model(self.h, model_id, pp, 8)

    def SetLinearDynamics(self, k, A, B, f=None):
        _, ap = _d(np.asarray(A).T)
        _, bp = _d(np.asarray(B).T)
        fp = _d(f)[1] if f is not None else None
        self.L.oracle_set_linear_dynamics(self.h, k, ap, bp, fp)

    def SetDiagonalCost(self, k, Qd, Rd, q, r, c):
        Rd = np.zeros(self.m) if Rd is None else Rd
        r = np.zeros(self.m) if r is None else r
        self.L.oracle_set_diagonal_cost(self.h, k, _d(Qd)[1], _d(Rd)[1], _d(q)[1], _d(r)[1],
                                        C.c_double(c))

    def SetQuadraticCost(self, k, Q, R, H, q, r, c):
        self.L.oracle_set_quadratic_cost(self.h, k, _d(np.asarray(Q).T)[1], _d(np.asarray(R).T)[1],
                                         _d(np.asarray(H).T)[1], _d(q)[1], _d(r)[1], C.c_double(c))

    def SetLQRCost(self, k, Qd, Rd, xref, uref):
        self.L.oracle_set_lqr_cost(self.h, k, _d(Qd)[1], _d(Rd)[1], _d(xref)[1], _d(uref)[1])

    def AddSelectorConstraint(self, k, cone, idx, scale, off):
        dim = len(idx)
        ia = (C.c_int * dim)(*[int(i) for i in idx])
        return self.L.oracle_add_constraint_selector(self.h, k, cone, dim, ia, _d(scale)[1],
                                                     _d(off)[1])

    def AddCallbackConstraint(self, k, cone, dim, con, jac):
        """con(x,u)->c[dim]; jac(x,u)->J[dim,n+m]"""
        n, m = self.n, self.m

        def c_con(_ud, out, xp, up):
            x = np.ctypeslib.as_array(xp, (n,))
            u = np.ctypeslib.as_array(up, (m,))
            v = np.asarray(con(x, u), dtype=float)
            for i in range(dim):
                out[i] = v[i]

        def c_jac(_ud, out, xp, up):
            x = np.ctypeslib.as_array(xp, (n,))
            u = np.ctypeslib.as_array(up, (m,))
            J = np.asarray(jac(x, u), dtype=float)
            flat = J.T.reshape(-1)
            for i in range(dim * (n + m)):
                out[i] = flat[i]

        cc, cj = CON_CB(c_con), CON_CB(c_jac)
        self._keep += [cc, cj]
        return self.L.oracle_add_constraint_callback(self.h, k, cone, dim, cc, cj, None)

    def SetInitialState(self, x0):
        self.L.oracle_set_initial_state(self.h, _d(x0)[1])

    def Initialize(self):
        return self.L.oracle_initialize(self.h)

    def SetState(self, x, k):
        self.L.oracle_set_state(self.h, k, _d(x)[1])

    def SetInput(self, u, k=None):
        ks = range(self.N) if k is None else [k]
        for kk in ks:
            self.L.oracle_set_input(self.h, kk, _d(u)[1])

    def SetDual(self, k, j, z):
        self.L.oracle_set_dual(self.h, k, j, _d(z)[1])

    def SetPenalty(self, rho):
        self.L.oracle_set_penalty(self.h, rho)

    def UpdateLinearCosts(self, k, q, r, c):
        self.L.oracle_update_linear_costs(self.h, k, _d(q)[1] if q is not None else None,
                                          _d(r)[1] if r is not None else None, C.c_double(c))

    def ShiftTrajectory(self):
        self.L.oracle_shift_trajectory(self.h)

    # --- solve + stages
    def Solve(self):
        self.L.oracle_solve(self.h)
        return self.L.oracle_get_status(self.h)

    def GetIterations(self):
        return self.L.oracle_get_iterations(self.h)

    def GetMeritEvals(self):
        return self.L.oracle_get_merit_evals(self.h)

    def GetFinalPhi(self):
        return self.L.oracle_get_final_phi(self.h)

    def OpenLoopRollout(self):
        self.L.oracle_open_loop_rollout(self.h)

    def LinearRollout(self):
        self.L.oracle_linear_rollout(self.h)

    def CopyTrajectory(self):
        self.L.oracle_copy_trajectory(self.h)

    def CalcCost(self):
        return self.L.oracle_calc_cost(self.h)

    def CalcCostGradient(self):
        self.L.oracle_calc_cost_gradient(self.h)

    def CalcExpansions(self):
        self.L.oracle_calc_expansions(self.h)

    def BackwardPass(self):
        return self.L.oracle_backward_pass(self.h)

    def MeritFunction(self, alpha, want_derivative=True):
        phi, dphi = C.c_double(), C.c_double()
        self.L.oracle_merit_function(self.h, alpha, C.byref(phi),
                                     C.byref(dphi) if want_derivative else None)
        return (phi.value, dphi.value) if want_derivative else phi.value

    def ForwardPass(self):
        a = C.c_double()
        err = self.L.oracle_forward_pass(self.h, C.byref(a))
        return err, a.value

    def Stationarity(self):
        return self.L.oracle_stationarity(self.h)

    def Feasibility(self):
        return self.L.oracle_feasibility(self.h)

    def DualUpdate(self):
        self.L.oracle_dual_update(self.h)

    def PenaltyUpdate(self):
        self.L.oracle_penalty_update(self.h)

    def LsIters(self):
        return self.L.oracle_ls_iters(self.h)

    def MeritValues(self):
        v = [C.c_double() for _ in range(4)]
        self.L.oracle_get_merit_values(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def KnotOp(self, k, name):
        self.L.oracle_knot_op(self.h, k, OPS[name])

    def KnotCalcCost(self, k):
        return self.L.oracle_knot_calc_cost(self.h, k)

    def KnotCalcConstraintCosts(self, k):
        return self.L.oracle_knot_calc_constraint_costs(self.h, k)

    def KnotCalcViolations(self, k):
        return self.L.oracle_knot_calc_violations(self.h, k)

    def get(self, k, name, shape=None):
        """KnotPointData member by name; matrices returned as (rows, cols) numpy arrays."""
        buf = np.zeros(4096)
        cnt = self.L.oracle_get_field(self.h, k, name.encode(), buf.ctypes.data_as(dptr))
        if cnt < 0:
            raise KeyError(name)
        out = buf[:cnt].copy()
        if shape is not None:
            out = out.reshape(shape[1], shape[0]).T.copy()
        return out

    def set(self, k, name, val):
        v = np.asarray(val, dtype=float)
        if v.ndim == 2:
            v = v.T
        a, p = _d(v.reshape(-1))
        if self.L.oracle_set_field(self.h, k, name.encode(), p) < 0:
            raise KeyError(name)

    def GetState(self, k):
        return self.get(k, "x_")

    def GetInput(self, k):
        return self.get(k, "u_")

    def states(self):
        return np.stack([self.get(k, "x_") for k in range(self.N + 1)])

    def inputs(self):
        return np.stack([self.get(k, "u_") for k in range(self.N)])


def make_batch_spec(P):
    """altro_b200.problems.Problem -> (BatchSpec, keepalive list)."""
    keep = []

    def dp(a):
        if a is None:
            return None
        arr, p = _d(a)
        keep.append(arr)
        return p

    sp = BatchSpec()
    sp.N, sp.n, sp.m, sp.B = P.N, P.n, P.m, P.B
    sp.h = P.h
    sp.model_id = P.model_id
    for i in range(8):
        sp.model_params[i] = P.model_params[i] if i < len(P.model_params) else 0.0
    sp.Qd, sp.Rd = dp(P.Qd), dp(P.Rd)
    sp.ref_mode = P.ref_mode
    sp.q, sp.r, sp.c = dp(P.q), dp(P.r), dp(P.c)
    sp.xref, sp.uref = dp(P.xref), dp(P.uref)
    if P.offsets is not None:
        off = np.ascontiguousarray(P.offsets, dtype=np.int32)
        keep.append(off)
        sp.offsets = off.ctypes.data_as(iptr)
    sp.T = P.T
    sp.x0, sp.U0 = dp(P.x0), dp(P.U0)
    sp.U0_per_problem = int(P.U0_per_problem)
    sp.ncon = len(P.constraints)
    for j, cs in enumerate(P.constraints):
        c = sp.con[j]
        c.k_start, c.k_stop, c.cone, c.dim = cs.k_start, cs.k_stop, cs.cone, len(cs.idx)
        for i in range(len(cs.idx)):
            c.idx[i] = int(cs.idx[i])
            c.scale[i] = float(cs.scale[i])
            c.off[i] = float(cs.off[i])
        c.off_b = dp(cs.off_b) if cs.off_b is not None else None
    o = default_options()
    for k, v in P.options.items():
        setattr(o, k, v)
    sp.opts = o
    return sp, keep


def solve_batch(P, b0=0, b1=None, nthreads=None, want_y=False):
    """Solve problems [b0,b1) of Problem P on the CPU oracle.  Returns dict of arrays + seconds."""
    L = lib()
    b1 = P.B if b1 is None else b1
    nb = b1 - b0
    nthreads = L.oracle_max_threads() if nthreads is None else nthreads
    sp, keep = make_batch_spec(P)
    X = np.zeros((nb, P.N + 1, P.n))
    U = np.zeros((nb, P.N, P.m))
    Y = np.zeros((nb, P.N + 1, P.n)) if want_y else None
    status = np.zeros(nb, dtype=np.int32)
    iters = np.zeros(nb, dtype=np.int32)
    evals = np.zeros(nb, dtype=np.int64)
    cost = np.zeros(nb)
    stat = np.zeros(nb)
    feas = np.zeros(nb)
    res = BatchResult()
    res.X = X.ctypes.data_as(dptr)
    res.U = U.ctypes.data_as(dptr)
    res.Y = Y.ctypes.data_as(dptr) if want_y else None
    res.status = status.ctypes.data_as(iptr)
    res.iters = iters.ctypes.data_as(iptr)
    res.merit_evals = evals.ctypes.data_as(C.POINTER(C.c_long))
    res.cost = cost.ctypes.data_as(dptr)
    res.stat = stat.ctypes.data_as(dptr)
    res.feas = feas.ctypes.data_as(dptr)
    secs = L.oracle_batch_solve(C.byref(sp), b0, b1, nthreads, C.byref(res))
    return dict(X=X, U=U, Y=Y, status=status, iters=iters, merit_evals=evals, cost=cost,
                stat=stat, feas=feas, seconds=secs, threads=nthreads)
