"""Batched trajectory-optimisation problem descriptions (pure numpy, host side).

A `Problem` is the batch analogue of what a user of the reference builds with B separate
`altro::ALTROSolver` objects: horizon, dimensions, time step (float, typedefs.hpp:31-35), a
dynamics model id, an LQR-type cost (altro_solver.cpp:138-172), constraints, initial states and
the initial control guess.  The generators below are the configurations of BASELINE.json, with
the seeds / distributions fixed in SURVEY.md section 8(d).
"""
import json
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# ConstraintType, src/altro/solver/typedefs.hpp:53
EQUALITY, IDENTITY, INEQUALITY, SECOND_ORDER_CONE = 0, 1, 2, 3
# device model ids (altro_b200/csrc/models.cuh)
MODEL_LINEAR, MODEL_DOUBLE_INTEGRATOR, MODEL_PENDULUM, MODEL_BICYCLE4, MODEL_BICYCLE5, MODEL_CHAIN = range(6)
# how the LQR reference of each problem is given
REF_SHARED, REF_FULL, REF_GOAL, REF_WINDOW = 0, 1, 2, 3

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class ConstraintSpec:
    """Rows c_i = scale_i * [x;u][idx_i] + off_i (idx_i = -1: c_i = off_i) on knots [k_start,k_stop)."""
    k_start: int
    k_stop: int
    cone: int
    idx: List[int]
    scale: List[float]
    off: List[float]
    off_b: Optional[np.ndarray] = None  # per-problem offsets [B, dim]


@dataclass
class Problem:
    name: str
    N: int
    n: int
    m: int
    B: int
    h: float
    model_id: int
    model_params: List[float]
    Qd: np.ndarray            # [(N+1), n]
    Rd: np.ndarray            # [N, m]
    ref_mode: int
    x0: np.ndarray            # [B, n]
    U0: np.ndarray            # [N, m] or [B, N, m]
    U0_per_problem: bool = False
    q: Optional[np.ndarray] = None
    r: Optional[np.ndarray] = None
    c: Optional[np.ndarray] = None
    xref: Optional[np.ndarray] = None
    uref: Optional[np.ndarray] = None
    offsets: Optional[np.ndarray] = None
    T: int = 0
    constraints: List[ConstraintSpec] = field(default_factory=list)
    options: dict = field(default_factory=dict)

    def subset(self, b0, b1):
        """Problems [b0, b1) as a new Problem (views, no copies of shared data)."""
        import copy
        P = copy.copy(self)
        P.B = b1 - b0
        P.x0 = self.x0[b0:b1]
        if self.U0_per_problem:
            P.U0 = self.U0[b0:b1]
        if self.ref_mode == REF_FULL:
            P.q, P.r, P.c = self.q[b0:b1], self.r[b0:b1], self.c[b0:b1]
        if self.ref_mode == REF_GOAL:
            P.xref, P.uref = self.xref[b0:b1], self.uref[b0:b1]
        if self.ref_mode == REF_WINDOW:
            P.offsets = self.offsets[b0:b1]
        P.constraints = []
        for cs in self.constraints:
            cs2 = copy.copy(cs)
            if cs.off_b is not None:
                cs2.off_b = cs.off_b[b0:b1]
            P.constraints.append(cs2)
        return P

    def n_constraint_rows(self):
        return sum(len(c.idx) for c in self.constraints)


def _f32(x):
    return float(np.float32(x))


def double_integrator(N=10, tf=5.0, variant="unconstrained", B=1, x0=None):
    """test/double_integrator_test.cpp:66-492 (N=10, h=0.5f) and BASELINE C0 (N=50)."""
    n, m, dim = 4, 2, 2
    h = _f32(tf / N)
    Qd = np.ones((N + 1, n))
    Rd = np.full((N, m), 1e-2)
    if x0 is None:
        x0 = [1.0, 2.0, 0.0, 0.0] if variant in ("unconstrained", "goal") else [2.0, 2.0, 0.0, 0.0]
    cons, opts = [], {}
    if variant in ("goal", "ubox", "usoc"):  # :174-190 goal constraint c = x - xf, terminal
        cons.append(ConstraintSpec(N, N + 1, EQUALITY, list(range(n)), [1.0] * n, [0.0] * n))
        opts["penalty_scaling"] = 100.0
    if variant == "ubox":  # :283-304 control box, dim 2m, knots 0..N-1
        cons.append(ConstraintSpec(0, N, INEQUALITY, [n, n + 1, n, n + 1], [1, 1, -1, -1],
                                   [-1.0, -1.0, -1.0, -1.0]))
        opts["penalty_initial"] = 100.0
    if variant == "usoc":  # :405-424 ||u|| <= 1, c = [u; 1]
        cons.append(ConstraintSpec(0, N, SECOND_ORDER_CONE, [n, n + 1, -1], [1, 1, 0],
                                   [0.0, 0.0, 1.0]))
        opts["penalty_initial"] = 1.0
    if variant == "unconstrained":
        opts["iterations_max"] = 3
    return Problem(name=f"double_integrator_{variant}_N{N}", N=N, n=n, m=m, B=B, h=h,
                   model_id=MODEL_DOUBLE_INTEGRATOR, model_params=[dim], Qd=Qd, Rd=Rd,
                   ref_mode=REF_GOAL, xref=np.zeros((B, n)), uref=np.zeros((B, m)),
                   x0=np.tile(np.asarray(x0, dtype=float), (B, 1)), U0=np.zeros((N, m)),
                   constraints=cons, options=opts)


def pendulum(B=4096, N=100, tf=3.0, goal_constraint=False, seed=0, perturb=True,
             iterations_max=100):
    """BASELINE C1; N=50/tf=3 and N=20/tf=2 are test/pendulum_test.cpp:45-203."""
    n, m = 2, 1
    h = _f32(tf / N)
    Qd = np.full((N + 1, n), 1e-2)
    Qd[N] = 1.0
    Rd = np.full((N, m), 1e-3)
    xf = np.array([math.pi, 0.0])
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-0.5, 0.5, size=(B, n)) if perturb else np.zeros((B, n))
    cons = []
    if goal_constraint:  # pendulum_test.cpp:160-173: c = xf - x at the terminal knot
        cons.append(ConstraintSpec(N, N + 1, EQUALITY, [0, 1], [-1.0, -1.0], list(xf)))
    return Problem(name=f"pendulum_N{N}" + ("_goal" if goal_constraint else ""), N=N, n=n, m=m,
                   B=B, h=h, model_id=MODEL_PENDULUM, model_params=[], Qd=Qd, Rd=Rd,
                   ref_mode=REF_GOAL, xref=np.tile(xf, (B, 1)), uref=np.zeros((B, m)), x0=x0,
                   U0=np.full((N, m), 0.1), constraints=cons,
                   options={"iterations_max": iterations_max})


def bicycle_turn90():
    """test/bicycle_test.cpp:53-138 (n=4, N=30, backtracking line search)."""
    n, m, N = 4, 2, 30
    h = _f32(3.0 / N)
    Qd = np.full((N + 1, n), 1e-2)
    Qd[N] = 10.0
    Rd = np.full((N, m), 1e-3)
    xf = np.array([1.0, 2.0, math.pi / 2, 0.0])
    return Problem(name="bicycle4_turn90", N=N, n=n, m=m, B=1, h=h, model_id=MODEL_BICYCLE4,
                   model_params=[2.7, 1.5], Qd=Qd, Rd=Rd, ref_mode=REF_GOAL, xref=xf[None],
                   uref=np.zeros((1, m)), x0=np.zeros((1, n)), U0=np.tile([0.5, 0.0], (N, 1)),
                   options={"iterations_max": 30, "use_backtracking_linesearch": 1})


def bicycle(B=16384, N=100, tf=3.0, seed=1, n=5, iterations_max=30):
    """BASELINE C2: random goals; n=5 model [x,y,theta,delta,v], u=[a,delta_dot].
    Options are those of the reference's own bicycle solve (bicycle_test.cpp:124-129):
    iterations_max = 30, backtracking line search."""
    m = 2
    h = _f32(tf / N)
    Qd = np.full((N + 1, n), 1e-2)
    Qd[N] = 10.0
    Rd = np.full((N, m), 1e-3)
    rng = np.random.default_rng(seed)
    px = rng.uniform(0.5, 2.0, B)
    py = rng.uniform(-2.0, 2.0, B)
    th = rng.uniform(-math.pi / 2, math.pi / 2, B)
    xf = np.zeros((B, n))
    xf[:, 0], xf[:, 1], xf[:, 2] = px, py, th
    x0 = np.zeros((B, n))
    if n == 5:
        x0[:, 4] = 0.5
        U0 = np.zeros((N, m))
        model = MODEL_BICYCLE5
    else:
        U0 = np.tile([0.5, 0.0], (N, 1))
        model = MODEL_BICYCLE4
    return Problem(name=f"bicycle{n}_N{N}", N=N, n=n, m=m, B=B, h=h, model_id=model,
                   model_params=[2.7, 1.5], Qd=Qd, Rd=Rd, ref_mode=REF_GOAL, xref=xf,
                   uref=np.zeros((B, m)), x0=x0, U0=U0,
                   options={"iterations_max": iterations_max, "use_backtracking_linesearch": 1})


def load_scotty():
    """altro_b200/data/scotty_ref.json (the numbers of test/scotty.json; input data of the tracking
    configs, written by tests/golden/make_golden.py): xref [501,4], uref [501,2], h."""
    d = json.load(open(os.path.join(_DATA, "scotty_ref.json")))
    xref = np.array(d["state_trajectory"], dtype=float)
    uref = np.array(d["input_trajectory"], dtype=float)
    Nref = int(d["N"]) - 1                    # test_utils.cpp:287
    tref = np.float32(d["tf"])                # float t_ref (bicycle_test.cpp:176)
    h = _f32(float(tref) / float(Nref))       # bicycle_test.cpp:180
    return xref, uref, h


def scotty(B=65536, N=50, n=5, seed=2, iterations_max=80, margin=0):
    """BASELINE C3: one tracking-MPC horizon per problem, windows along the scotty trajectory,
    steering-angle bound +-60deg at every knot (test/bicycle_test.cpp:144-224).
    margin: rows of the reference table left free behind every window, i.e. how many
    receding-horizon steps each problem can take."""
    m = 2
    xr4, ur4, h = load_scotty()
    T = xr4.shape[0]
    if n == 5:
        xtab = np.concatenate([xr4, ur4[:, :1]], axis=1)
        utab = np.stack([np.zeros(T), ur4[:, 1]], axis=1)
        sigma = np.array([0.05, 0.05, 0.02, 0.01, 0.05])
        model = MODEL_BICYCLE5
    else:
        xtab, utab = xr4, ur4
        sigma = np.array([0.05, 0.05, 0.02, 0.01])
        model = MODEL_BICYCLE4
    offsets = (np.arange(B) % (T - 1 - N - margin)).astype(np.int32)
    rng = np.random.default_rng(seed)
    x0 = xtab[offsets] + rng.normal(size=(B, n)) * sigma
    U0 = np.zeros((B, N, m))
    if n == 4:
        U0[:, :, 0] = ur4[offsets, 0][:, None]
    Qd = np.full((N + 1, n), 1e-2)
    Rd = np.full((N, m), 1e-3)
    dmax = 60 * math.pi / 180.0
    cons = [ConstraintSpec(0, N + 1, INEQUALITY, [3, 3], [1.0, -1.0], [-dmax, -dmax])]
    return Problem(name=f"scotty{n}_N{N}", N=N, n=n, m=m, B=B, h=h, model_id=model,
                   model_params=[2.7, 1.5], Qd=Qd, Rd=Rd, ref_mode=REF_WINDOW, xref=xtab,
                   uref=utab, offsets=offsets, T=T, x0=x0, U0=U0, U0_per_problem=True,
                   constraints=cons,
                   options={"iterations_max": iterations_max, "use_backtracking_linesearch": 1})


def chain(B=32768, n=4, m=2, N=50, seed=3, control_box=False, iterations_max=100, hard=False):
    """BASELINE C4 dimension sweep: coupled pendulum chain, h=0.02f.  hard: initial
    states three pendulum-radians out with a stiffer terminal weight -- the iteration needs 5-20
    passes instead of the two of SURVEY 8d's near-linear instances, so a timing of the sweep
    measures the loop rather than the prologue."""
    h = _f32(0.02)
    Qd = np.full((N + 1, n), 1e-2)
    Qd[N] = 10.0 if hard else 1.0
    Rd = np.full((N, m), 1e-2 if hard else 1e-3)
    rng = np.random.default_rng(seed)
    amp = 3.0 if hard else 0.5
    x0 = rng.uniform(-amp, amp, size=(B, n))
    cons = []
    if control_box:
        ub = 2.0
        idx = [n + i for i in range(m)] * 2
        cons.append(ConstraintSpec(0, N, INEQUALITY, idx, [1.0] * m + [-1.0] * m, [-ub] * (2 * m)))
    return Problem(name=f"chain_n{n}_m{m}_N{N}" + ("_ubox" if control_box else "") + ("_hard" if hard else ""),
                   N=N, n=n, m=m,
                   B=B, h=h, model_id=MODEL_CHAIN, model_params=[n, m], Qd=Qd, Rd=Rd,
                   ref_mode=REF_GOAL, xref=np.zeros((B, n)), uref=np.zeros((B, m)), x0=x0,
                   U0=np.full((N, m), 0.05), constraints=cons,
                   options={"iterations_max": iterations_max})
