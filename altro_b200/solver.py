"""ctypes binding of libaltro_b200.so; method names follow altro::ALTROSolver.

Every call goes through the C ABI (include/altro_b200.h) -- the same entry points a cgo / JNI /
C++ consumer of the reference would bind.  numpy arrays are host buffers in the
problem-major layout [B][...].
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libaltro_b200.so")
_LIB = None

dptr = C.POINTER(C.c_double)
iptr = C.POINTER(C.c_int)


class ErrorCodes:  # src/altro/solver/exceptions.hpp:24-51
    NoError, StateDimUnknown, InputDimUnknown, NextStateDimUnknown, DimensionUnknown, BadIndex, \
        DimensionMismatch, SolverNotInitialized, SolverAlreadyInitialized, NonPositive, \
        TimestepNotPositive, CostFunNotSet, DynamicsFunNotSet, InvalidOptAtTerminalKnotPoint, \
        MaxConstraintsExceeded, InvalidConstraintDim, CholeskyFailed, OpOnlyValidAtTerminalKnotPoint, \
        InvalidPointer, BackwardPassFailed, LineSearchFailed, MeritFunctionGradientTooSmall, \
        InvalidBoundConstraint, NonPositivePenalty, CostNotQuadratic, FileError = range(26)
    NoDevice, Unsupported = 100, 101


class SolveStatus:  # src/altro/solver/typedefs.hpp:19-27
    Success, Unsolved, MaxIterations = 0, 1, 2


LastIndex, AllIndices = -1, -2


class Options(C.Structure):  # AltroOptions, solver_options.hpp:16-39
    _fields_ = [("iterations_max", C.c_int), ("tol_primal_feasibility", C.c_double),
                ("tol_stationarity", C.c_double), ("tol_meritfun_gradient", C.c_double),
                ("penalty_initial", C.c_double), ("penalty_scaling", C.c_double),
                ("penalty_max", C.c_double), ("use_backtracking_linesearch", C.c_int),
                ("linesearch_c1", C.c_double), ("linesearch_c2", C.c_double)]


class AltroB200Error(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        msg = load_library().altro_b200_error_string(code).decode()
        super().__init__(f"ALTRO ERROR Code {code}: {msg} {where}")


def build_library(verbose=False):
    from . import build as _build
    return _build.build(verbose=verbose)[0]


def load_library():
    """Loads the CUDA library.  Fails loudly when it has not been built (no fallback path)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            raise RuntimeError(f"{_LIBPATH} is missing: run `python -m altro_b200.build` "
                               "(the solve path is CUDA-only; there is no CPU fallback)")
        L = C.CDLL(_LIBPATH)
        L.altro_b200_error_string.restype = C.c_char_p
        L.altro_b200_create.restype = C.c_void_p
        L.altro_b200_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.altro_b200_destroy.argtypes = [C.c_void_p]
        L.altro_b200_kernel_launches.restype = C.c_long
        L.altro_b200_kernel_launches.argtypes = [C.c_void_p]
        L.altro_b200_device_bytes.restype = C.c_long
        L.altro_b200_device_bytes.argtypes = [C.c_void_p]
        L.altro_b200_set_time_step.argtypes = [C.c_void_p, C.c_float]
        L.altro_b200_set_time_step_range.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        L.altro_b200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        vp = C.c_void_p
        L.altro_b200_set_dimension.argtypes = [vp, C.c_int, C.c_int]
        L.altro_b200_set_model.argtypes = [vp, C.c_int, dptr, C.c_int]
        L.altro_b200_set_linear_dynamics.argtypes = [vp, dptr, dptr, dptr, C.c_int, C.c_int]
        L.altro_b200_set_lqr_cost.argtypes = [vp, dptr, dptr, dptr, dptr, C.c_int, C.c_int, C.c_int]
        L.altro_b200_set_lqr_cost_window.argtypes = [vp, dptr, dptr, dptr, dptr, C.c_int, iptr]
        L.altro_b200_set_diagonal_cost.argtypes = [vp, dptr, dptr, dptr, dptr, dptr, C.c_int, C.c_int, C.c_int]
        L.altro_b200_update_linear_costs.argtypes = [vp, dptr, dptr, dptr, C.c_int, C.c_int, C.c_int]
        L.altro_b200_set_quadratic_cost.argtypes = [vp, dptr, dptr, dptr, dptr, dptr, dptr, C.c_int, C.c_int, C.c_int]
        L.altro_b200_set_constraint_affine.argtypes = [vp, C.c_int, C.c_int, dptr, dptr, dptr, C.c_int, C.c_int]
        L.altro_b200_set_constraint_disc.argtypes = [vp, C.c_int, C.c_int, dptr, dptr, C.c_int, C.c_int]
        L.altro_b200_advance_window.argtypes = [vp, C.c_int]
        L.altro_b200_advance_window_linear.argtypes = [vp, C.c_int, C.c_double]
        L.altro_b200_set_mpc_cost_update.argtypes = [vp, C.c_int, C.c_double]
        L.altro_b200_set_constraint.argtypes = [vp, C.c_int, C.c_int, iptr, dptr, dptr, dptr, C.c_int, C.c_int]
        L.altro_b200_set_initial_state.argtypes = [vp, dptr, C.c_int]
        L.altro_b200_initialize.argtypes = [vp]
        L.altro_b200_set_input.argtypes = [vp, dptr, C.c_int, C.c_int, C.c_int]
        L.altro_b200_set_state.argtypes = [vp, dptr, C.c_int, C.c_int, C.c_int]
        L.altro_b200_set_options.argtypes = [vp, C.POINTER(Options)]
        L.altro_b200_calc_cost.argtypes = [vp, dptr]
        L.altro_b200_set_penalty.argtypes = [vp, C.c_double]
        for f in ("reset_duals", "reset_trajectory", "shift_trajectory", "solve", "solve_async", "synchronize",
                  "open_loop_rollout", "mpc_step", "knot_eval"):
            getattr(L, "altro_b200_" + f).argtypes = [vp]
        for f in ("get_states", "get_inputs", "get_dual_dynamics", "get_feedback_gains",
                  "get_feedforward_gains", "get_final_objective", "get_stationarity",
                  "get_primal_feasibility", "get_penalty"):
            getattr(L, "altro_b200_" + f).argtypes = [vp, dptr]
        for f in ("get_status", "get_iterations", "get_merit_evals"):
            getattr(L, "altro_b200_" + f).argtypes = [vp, iptr]
        L.altro_b200_get_linesearch_histogram.argtypes = [vp, C.POINTER(C.c_long), C.c_int]
        L.altro_b200_get_field.argtypes = [vp, C.c_char_p, dptr, C.POINTER(C.c_int)]
        L.altro_b200_get_dual_general.argtypes = [vp, C.c_int, C.c_int, dptr]
        L.altro_b200_set_dual_general.argtypes = [vp, C.c_int, C.c_int, dptr, C.c_int]
        L.altro_b200_get_num_constraints.argtypes = [vp]
        L.altro_b200_set_solve_mode.argtypes = [vp, C.c_int]
        L.altro_b200_set_backward_mode.argtypes = [vp, C.c_int]
        L.altro_b200_set_profiling.argtypes = [vp, C.c_int]
        L.altro_b200_set_speculation.argtypes = [vp, C.c_int]
        L.altro_b200_set_candidate_store.argtypes = [vp, C.c_int]
        L.altro_b200_set_pipeline_split.argtypes = [vp, C.c_int]
        L.altro_b200_get_phase_stats.argtypes = [vp, dptr, C.POINTER(C.c_long), dptr, C.POINTER(C.c_long)]
        L.altro_b200_tvlqr_backward_batch.argtypes = [C.c_int] * 4 + [dptr] * 8 + [C.c_double, C.c_bool] + \
            [dptr] * 5 + [iptr]
        L.altro_b200_tvlqr_forward_batch.argtypes = [C.c_int] * 4 + [dptr] * 11
        _LIB = L
    return _LIB


def device_count():
    return load_library().altro_b200_device_count()


def default_options(**kw):
    o = Options()
    load_library().altro_b200_default_options(C.byref(o))
    for k, v in kw.items():
        if k in ("ls_c1", "ls_c2"):
            k = "linesearch_" + k[3:]
        setattr(o, k, v)
    return o


def _d(a):
    """contiguous float64 array + pointer (None passes through as NULL)"""
    if a is None:
        return None, None
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return arr, arr.ctypes.data_as(dptr)


class BatchSolver:
    """B independent `altro::ALTROSolver(N)` problems on one GPU."""

    def __init__(self, horizon_length, batch=1, device=0):
        self.L = load_library()
        self.N, self.B = horizon_length, batch
        self.n = self.m = 0
        self.h = self.L.altro_b200_create(horizon_length, batch, device)
        if not self.h:
            raise AltroB200Error(ErrorCodes.NoDevice, "(altro_b200_create)")
        self.h = C.c_void_p(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.altro_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, code, where):
        if code != 0:
            raise AltroB200Error(code, f"({where})")
        return code

    # ---- problem definition (altro_solver.hpp:37-330)
    def SetDimension(self, num_states, num_inputs):
        self._ck(self.L.altro_b200_set_dimension(self.h, num_states, num_inputs), "SetDimension")
        self.n, self.m = num_states, num_inputs

    def SetTimeStep(self, h, k_start=None, k_stop=0):
        """SetTimeStep(h) for the whole horizon, or (h, k_start, k_stop) for knots [k_start, k_stop)."""
        if k_start is None:
            self._ck(self.L.altro_b200_set_time_step(self.h, C.c_float(h)), "SetTimeStep")
        else:
            self._ck(self.L.altro_b200_set_time_step_range(self.h, C.c_float(h), k_start, k_stop), "SetTimeStep")

    def SetExplicitDynamics(self, model_id, params=()):
        prm, pp = _d(list(params) + [0.0] * (8 - len(params)))
        self._ck(self.L.altro_b200_set_model(self.h, model_id, pp, 8), "SetExplicitDynamics")

    def SetLinearDynamics(self, A, B, f=None, k_start=AllIndices, k_stop=0):
        a, ap = _d(np.asarray(A, dtype=float).T)
        b, bp = _d(np.asarray(B, dtype=float).T)
        ff, fp = _d(f)
        self._ck(self.L.altro_b200_set_linear_dynamics(self.h, ap, bp, fp, k_start, k_stop),
                 "SetLinearDynamics")

    def SetLQRCost(self, Qd, Rd, xref, uref, k_start=AllIndices, k_stop=0):
        xr = np.asarray(xref, dtype=float)
        per = 1 if xr.ndim == 2 else 0
        a, ap = _d(Qd)
        b, bp = _d(Rd)
        x, xp = _d(xr)
        u, up = _d(uref)
        self._ck(self.L.altro_b200_set_lqr_cost(self.h, ap, bp, xp, up, per, k_start, k_stop), "SetLQRCost")

    def SetLQRCostWindow(self, Qd, Rd, xtab, utab, offsets):
        a, ap = _d(Qd)
        b, bp = _d(Rd)
        x, xp = _d(xtab)
        u, up = _d(utab)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        self._ck(self.L.altro_b200_set_lqr_cost_window(self.h, ap, bp, xp, up, x.shape[0],
                                                       off.ctypes.data_as(iptr)), "SetLQRCostWindow")

    def SetDiagonalCost(self, Qd, Rd, q, r, c, k_start=AllIndices, k_stop=0):
        qq = np.asarray(q, dtype=float)
        per = 1 if qq.ndim >= 2 else 0
        a, ap = _d(Qd)
        b, bp = _d(Rd)
        qa, qp = _d(qq)
        ra, rp = _d(r)
        ca, cp = _d(np.atleast_1d(np.asarray(c, dtype=float)))
        self._ck(self.L.altro_b200_set_diagonal_cost(self.h, ap, bp, qp, rp, cp, per, k_start, k_stop),
                 "SetDiagonalCost")

    def SetQuadraticCost(self, Q, R, H, q, r, c, k_start=AllIndices, k_stop=0):
        """Dense Q (n x n), R (m x m), H (m x n) as 2-D arrays (stored column-major for the ABI)."""
        qq = np.asarray(q, dtype=float)
        per = 1 if qq.ndim >= 2 else 0
        Qa, Qp = _d(np.asarray(Q, dtype=float).T)
        Ra, Rp = _d(None if R is None else np.asarray(R, dtype=float).T)
        Ha, Hp = _d(None if H is None else np.asarray(H, dtype=float).T)
        qa, qp = _d(qq)
        ra, rp = _d(r)
        ca, cp = _d(np.atleast_1d(np.asarray(c, dtype=float)))
        self._ck(self.L.altro_b200_set_quadratic_cost(self.h, Qp, Rp, Hp, qp, rp, cp, per, k_start, k_stop),
                 "SetQuadraticCost")

    def UpdateLinearCosts(self, q, r, c, k_start=AllIndices, k_stop=0):
        qa, qp = _d(q)
        ra, rp = _d(r)
        ca, cp = _d(np.atleast_1d(np.asarray(c, dtype=float)))
        per = 1 if (qa is not None and qa.ndim >= 2) or ca.size > 1 else 0
        self._ck(self.L.altro_b200_update_linear_costs(self.h, qp, rp, cp, per, k_start, k_stop),
                 "UpdateLinearCosts")

    def AdvanceWindow(self, steps=1):
        self._ck(self.L.altro_b200_advance_window(self.h, steps), "AdvanceWindow")

    def AdvanceWindowLinear(self, steps, c_u):
        """The reference's MPC cost update (UpdateLinearCosts(q, nullptr, c) on the moved window)."""
        self._ck(self.L.altro_b200_advance_window_linear(self.h, steps, C.c_double(c_u)), "AdvanceWindowLinear")

    def SetMpcCostUpdate(self, mode, c_u=0.0):
        self._ck(self.L.altro_b200_set_mpc_cost_update(self.h, mode, C.c_double(c_u)), "SetMpcCostUpdate")

    def SetConstraint(self, cone, idx, scale, off, k_start, k_stop=0, off_b=None):
        dim = len(idx)
        ia = (C.c_int * dim)(*[int(i) for i in idx])
        s, sp = _d(scale)
        o, op = _d(off)
        ob, obp = _d(off_b)
        self._ck(self.L.altro_b200_set_constraint(self.h, cone, dim, ia, sp, op, obp, k_start, k_stop),
                 "SetConstraint")

    def SetConstraintAffine(self, cone, J, e, k_start, k_stop=0, e_b=None):
        """General affine rows c = J [x;u] + e, J a 2-D array (dim x (n+m))."""
        Jm = np.asarray(J, dtype=float)
        Ja, Jp = _d(Jm.T)  # column-major
        ea, ep = _d(e)
        eb, ebp = _d(e_b)
        self._ck(self.L.altro_b200_set_constraint_affine(self.h, cone, Jm.shape[0], Jp, ep, ebp, k_start, k_stop),
                 "SetConstraintAffine")

    def SetConstraintDisc(self, idx_a, idx_b, disc, k_start, k_stop=0, disc_b=None):
        """Keep-out disc r^2 - (v_a-cx)^2 - (v_b-cy)^2 <= 0; disc = (cx, cy, r)."""
        da, dp = _d(disc)
        db, dbp = _d(disc_b)
        self._ck(self.L.altro_b200_set_constraint_disc(self.h, idx_a, idx_b, dp, dbp, k_start, k_stop),
                 "SetConstraintDisc")

    def SetInitialState(self, x0):
        x = np.asarray(x0, dtype=float)
        a, ap = _d(x)
        self._ck(self.L.altro_b200_set_initial_state(self.h, ap, 1 if x.ndim == 2 else 0), "SetInitialState")

    def Initialize(self):
        self._ck(self.L.altro_b200_initialize(self.h), "Initialize")

    def SetInput(self, u, k_start=AllIndices, k_stop=0):
        a = np.asarray(u, dtype=float)
        layout = {1: 0, 2: 1, 3: 2}[a.ndim]
        arr, p = _d(a)
        self._ck(self.L.altro_b200_set_input(self.h, p, layout, k_start, k_stop), "SetInput")

    def SetState(self, x, k_start=AllIndices, k_stop=0):
        a = np.asarray(x, dtype=float)
        layout = {1: 0, 2: 1, 3: 2}[a.ndim]
        arr, p = _d(a)
        self._ck(self.L.altro_b200_set_state(self.h, p, layout, k_start, k_stop), "SetState")

    def SetOptions(self, opts):
        self._ck(self.L.altro_b200_set_options(self.h, C.byref(opts)), "SetOptions")

    def SetStream(self, cuda_stream_ptr):
        self._ck(self.L.altro_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "SetStream")

    def SetSolveMode(self, mode):
        """0: two-kernel-per-iteration pipeline (default); 1: single persistent kernel (test twin)."""
        self._ck(self.L.altro_b200_set_solve_mode(self.h, mode), "SetSolveMode")

    def SetBackwardMode(self, team):
        """0: Riccati sweep by one warp per group; 1: by the warps of a CTA (column teams)."""
        self._ck(self.L.altro_b200_set_backward_mode(self.h, int(team)), "SetBackwardMode")

    def SetSpeculation(self, nslots):
        self._ck(self.L.altro_b200_set_speculation(self.h, nslots), "SetSpeculation")

    def SetCandidateStore(self, nstore):
        self._ck(self.L.altro_b200_set_candidate_store(self.h, nstore), "SetCandidateStore")

    def SetPipelineSplit(self, nsplit):
        """0: automatic; 1..8 sub-batches pipelined on their own streams."""
        self._ck(self.L.altro_b200_set_pipeline_split(self.h, nsplit), "SetPipelineSplit")

    def SetProfiling(self, on):
        self._ck(self.L.altro_b200_set_profiling(self.h, int(on)), "SetProfiling")

    PHASES = ("init_rollout", "expand", "backward", "forward", "fwd_rollout", "fwd_expand",
              "fwd_dphi_ls", "fwd_criteria")

    def GetPhaseStats(self):
        ms = np.zeros(8)
        units = np.zeros(8)
        launches = (C.c_long * 8)()
        syncs = C.c_long()
        self._ck(self.L.altro_b200_get_phase_stats(self.h, ms.ctypes.data_as(dptr), launches,
                                                   units.ctypes.data_as(dptr), C.byref(syncs)), "GetPhaseStats")
        return {p: dict(ms=float(ms[i]), launches=int(launches[i]), units=float(units[i]))
                for i, p in enumerate(self.PHASES)}, int(syncs.value)

    def GetLinesearchHistogram(self, reset=False):
        h = (C.c_long * 32)()
        self._ck(self.L.altro_b200_get_linesearch_histogram(self.h, h, int(reset)), "GetLinesearchHistogram")
        return np.array(list(h))

    def ResetDuals(self):
        self._ck(self.L.altro_b200_reset_duals(self.h), "ResetDuals")

    def ResetTrajectory(self):
        self._ck(self.L.altro_b200_reset_trajectory(self.h), "ResetTrajectory")

    def MpcStep(self):
        """x0 <- x_[1], ShiftTrajectory, advance the tracking window: one MPC step on the device."""
        self._ck(self.L.altro_b200_mpc_step(self.h), "MpcStep")

    def ShiftTrajectory(self):
        self._ck(self.L.altro_b200_shift_trajectory(self.h), "ShiftTrajectory")

    def KnotEval(self):
        """Per-knot expansions (A, B, constraint values, projected duals, lx, lu) at x_, u_."""
        self._ck(self.L.altro_b200_knot_eval(self.h), "KnotEval")

    def SetPenalty(self, rho):
        self._ck(self.L.altro_b200_set_penalty(self.h, C.c_double(rho)), "SetPenalty")

    def OpenLoopRollout(self):
        self._ck(self.L.altro_b200_open_loop_rollout(self.h), "OpenLoopRollout")

    def CalcCost(self):
        return self._get("calc_cost", (self.B,))

    # ---- solve
    def Solve(self):
        self._ck(self.L.altro_b200_solve(self.h), "Solve")
        return self.GetStatus()

    def SolveAsync(self):
        self._ck(self.L.altro_b200_solve_async(self.h), "Solve")

    def Synchronize(self):
        self._ck(self.L.altro_b200_synchronize(self.h), "Synchronize")

    def KernelLaunches(self):
        return self.L.altro_b200_kernel_launches(self.h)

    def DeviceBytes(self):
        return self.L.altro_b200_device_bytes(self.h)

    # ---- getters
    def _get(self, fn, shape, dtype=np.float64, out=None):
        if out is None:
            out = np.empty(shape, dtype=dtype)
        ptr = out.ctypes.data_as(dptr if dtype == np.float64 else iptr)
        self._ck(getattr(self.L, "altro_b200_" + fn)(self.h, ptr), fn)
        return out

    def GetStates(self, out=None):
        return self._get("get_states", (self.B, self.N + 1, self.n), out=out)

    def GetInputs(self, out=None):
        return self._get("get_inputs", (self.B, self.N, self.m), out=out)

    def GetDualDynamics(self):
        return self._get("get_dual_dynamics", (self.B, self.N + 1, self.n))

    def GetFeedbackGains(self):
        """[B, N, m, n]"""
        K = self._get("get_feedback_gains", (self.B, self.N, self.n, self.m))
        return np.swapaxes(K, 2, 3)

    def GetFeedforwardGains(self):
        return self._get("get_feedforward_gains", (self.B, self.N, self.m))

    def GetField(self, name):
        """KnotPointData member `name` of every problem: array [B, N+1, rows]."""
        rows = C.c_int()
        self._ck(self.L.altro_b200_get_field(self.h, name.encode(), None, C.byref(rows)), "GetField")
        out = np.empty((self.B, self.N + 1, rows.value))
        self._ck(self.L.altro_b200_get_field(self.h, name.encode(), out.ctypes.data_as(dptr), None), "GetField")
        return out

    def GetDualGeneral(self, constraint, k, dim):
        out = np.empty((self.B, dim))
        self._ck(self.L.altro_b200_get_dual_general(self.h, constraint, k, out.ctypes.data_as(dptr)), "GetDualGeneral")
        return out

    def SetDualGeneric(self, constraint, k, z):
        za = np.asarray(z, dtype=float)
        a, p = _d(za)
        self._ck(self.L.altro_b200_set_dual_general(self.h, constraint, k, p, 1 if za.ndim == 2 else 0), "SetDualGeneric")

    def GetStatus(self, out=None):
        return self._get("get_status", (self.B,), np.int32, out=out)

    def GetIterations(self):
        return self._get("get_iterations", (self.B,), np.int32)

    def GetMeritEvals(self):
        return self._get("get_merit_evals", (self.B,), np.int32)

    def GetFinalObjective(self, out=None):
        return self._get("get_final_objective", (self.B,), out=out)

    def GetStationarity(self):
        return self._get("get_stationarity", (self.B,))

    def GetPrimalFeasibility(self):
        return self._get("get_primal_feasibility", (self.B,))

    def GetPenalty(self):
        return self._get("get_penalty", (self.B,))


def make_solver(P, device=0, nslots=None):
    """Builds and initialises a BatchSolver from a problems.Problem (the same sequence of calls
    the reference tests make: SetDimension, SetTimeStep, SetExplicitDynamics, SetLQRCost,
    SetConstraint, SetInitialState, Initialize, SetInput)."""
    from . import problems as PR
    s = BatchSolver(P.N, P.B, device)
    s.SetDimension(P.n, P.m)
    s.SetTimeStep(P.h)
    s.SetExplicitDynamics(P.model_id, P.model_params)
    set_cost(s, P)
    for cs in P.constraints:
        s.SetConstraint(cs.cone, cs.idx, cs.scale, cs.off, cs.k_start, cs.k_stop, off_b=cs.off_b)
    s.SetInitialState(P.x0)
    if nslots is not None:
        s.SetSpeculation(nslots)
    s.Initialize()
    s.SetInput(P.U0)
    s.SetOptions(default_options(**P.options))
    return s


def set_cost(s, P):
    from . import problems as PR
    N = P.N
    if P.ref_mode == PR.REF_WINDOW:
        # weights are uniform over the horizon in the tracking configs
        s.SetLQRCostWindow(P.Qd[0], P.Rd[0], P.xref, P.uref, P.offsets)
    elif P.ref_mode == PR.REF_GOAL:
        # runs of knots with identical weights -> one SetLQRCost call each (as the tests do)
        k = 0
        while k <= N:
            k2 = k + 1
            while k2 <= N and np.array_equal(P.Qd[k2], P.Qd[k]) and \
                    (k2 >= N or np.array_equal(P.Rd[k2], P.Rd[min(k, N - 1)])):
                k2 += 1
            s.SetLQRCost(P.Qd[k], P.Rd[min(k, N - 1)], P.xref, P.uref, k, k2)
            k = k2
    elif P.ref_mode == PR.REF_SHARED:
        for k in range(N + 1):
            s.SetDiagonalCost(P.Qd[k], P.Rd[min(k, N - 1)], P.q[k], P.r[min(k, N - 1)], P.c[k], k, k + 1)
    else:  # REF_FULL
        for k in range(N + 1):
            s.SetDiagonalCost(P.Qd[k], P.Rd[min(k, N - 1)], P.q[:, k:k + 1, :],
                              P.r[:, min(k, N - 1):min(k, N - 1) + 1, :], P.c[:, k:k + 1], k, k + 1)


def solve_problem(P, device=0, mode=0):
    """Convenience: build, solve, gather.  Returns the same dict as oracle.solve_batch."""
    s = make_solver(P, device)
    s.SetSolveMode(mode)
    status = s.Solve()
    out = dict(X=s.GetStates(), U=s.GetInputs(), Y=s.GetDualDynamics(), status=status,
               iters=s.GetIterations(), merit_evals=s.GetMeritEvals(), cost=s.GetFinalObjective(),
               stat=s.GetStationarity(), feas=s.GetPrimalFeasibility())
    s.close()
    return out
