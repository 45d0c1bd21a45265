"""Parity tests proper: the CUDA solve path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Run on the B200 box: python -m pytest tests -m gpu"""
import math

import numpy as np
import pytest

import altro_b200
from altro_b200 import problems as PR
from parity_util import compare

pytestmark = pytest.mark.gpu


def run_both(oracle, P, nref=None):
    gpu = altro_b200.solve_problem(P)
    ref = oracle.solve_batch(P, 0, nref or P.B)
    return gpu, ref


@pytest.mark.parametrize("variant,iters", [("unconstrained", 1), ("goal", 3), ("ubox", 5), ("usoc", 9)])
def test_double_integrator_variants(oracle, variant, iters):
    """double_integrator_test.cpp: pinned iteration counts 3 / 5 / 9 (and 1 unconstrained)."""
    P = PR.double_integrator(N=10, variant=variant, B=3)
    gpu, ref = run_both(oracle, P)
    rep = compare(gpu, ref, tail_frac=0.0)
    assert np.all(gpu["iters"] == iters) and np.all(gpu["status"] == 0)
    assert np.linalg.norm(gpu["X"][0, -1]) < (1e-1 if variant == "unconstrained" else 1e-4)


def test_double_integrator_n50(oracle):
    """BASELINE config C0 shape (N=50, h=0.1f), all four variants."""
    for variant in ("unconstrained", "goal", "ubox", "usoc"):
        P = PR.double_integrator(N=50, variant=variant, B=2)
        P.options.pop("iterations_max", None)
        compare(*run_both(oracle, P))


def test_pendulum_reference_cases(oracle):
    """pendulum_test.cpp: unconstrained N=50 golden x_N, goal-constrained N=20."""
    P = PR.pendulum(B=1, N=50, tf=3.0, perturb=False, iterations_max=20)
    gpu, ref = run_both(oracle, P)
    compare(gpu, ref)
    assert gpu["status"][0] == 0 and gpu["iters"][0] == 10
    assert np.linalg.norm(gpu["X"][0, -1] - [3.12099917161669, 0.0011966258762942175]) < 1e-5
    P = PR.pendulum(B=1, N=20, tf=2.0, perturb=False, goal_constraint=True, iterations_max=100)
    gpu, ref = run_both(oracle, P)
    compare(gpu, ref)
    assert gpu["status"][0] == 0 and gpu["iters"][0] == 9
    assert np.linalg.norm(gpu["X"][0, -1] - [math.pi, 0]) < 1e-4


@pytest.mark.parametrize("goal", [False, True])
def test_pendulum_batch(oracle, goal):
    """BASELINE C1 shape (N=100), perturbed initial states, ragged batch size."""
    P = PR.pendulum(B=200, N=100, goal_constraint=goal)
    rep = compare(*run_both(oracle, P))
    assert rep["n"] == 200


def test_bicycle_turn90(oracle):
    """bicycle_test.cpp:53-138 (n=4, N=30, backtracking)."""
    P = PR.bicycle_turn90()
    gpu, ref = run_both(oracle, P)
    compare(gpu, ref)
    assert gpu["iters"][0] == 17 and gpu["merit_evals"][0] == 40
    assert np.linalg.norm(gpu["X"][0, -1] - P.xref[0]) < 1e-2


@pytest.mark.parametrize("n", [4, 5])
def test_bicycle_batch(oracle, n):
    """BASELINE C2 shape: random goals, N=100."""
    P = PR.bicycle(B=160, N=100, n=n)
    rep = compare(*run_both(oracle, P))
    assert rep["strict"] >= 25


@pytest.mark.parametrize("n,N", [(4, 30), (5, 50)])
def test_scotty_batch(oracle, n, N):
    """BASELINE C3 shape: tracking windows with the steering-angle inequality at every knot."""
    P = PR.scotty(B=96, N=N, n=n)
    gpu, ref = run_both(oracle, P)
    rep = compare(gpu, ref)
    assert rep["converged"] > 0.7 * P.B


@pytest.mark.parametrize("n,m,N,box", [(4, 2, 50, False), (4, 4, 50, True), (6, 2, 50, True),
                                        (6, 4, 200, False), (12, 2, 50, False), (12, 4, 50, True)])
def test_chain_sweep(oracle, n, m, N, box):
    """BASELINE C4 dimension sweep family."""
    P = PR.chain(B=40, n=n, m=m, N=N, control_box=box)
    compare(*run_both(oracle, P))


@pytest.mark.parametrize("n,m,N", [(12, 2, 200), (12, 4, 500), (6, 2, 500), (4, 4, 500), (6, 4, 200)])
def test_chain_sweep_long_horizons_and_large_blocks(oracle, n, m, N):
    """The sweep shapes the round-1 tests never compared with the oracle (N = 500, n = 12 at
    N >= 200), on the hard instances (5-20 iterations) the sweep table is measured on."""
    P = PR.chain(B=33, n=n, m=m, N=N, hard=True)
    gpu, ref = run_both(oracle, P)
    rep = compare(gpu, ref, tail_frac=0.07)
    assert rep["converged"] >= 0.9 * P.B and rep["mean_iters"] > 4


def test_mpc_warm_start_sequence(oracle):
    """The reference's 200-step receding-horizon run (bicycle_test.cpp:247-337) on one handle:
    Solve -> GetInput -> simulate -> UpdateLinearCosts(q, nullptr, c) per knot -> SetInitialState
    -> ShiftTrajectory, duals and penalties carried over (quirks Q3, Q14).  Iteration counts must
    equal the golden test/scotty_mpc.json; closed-loop states agree to the solver tolerance."""
    import json, os
    xref, uref, h = PR.load_scotty()
    n, m, N = 4, 2, 30
    Qd, Rd = np.full(n, 1e-2), np.full(m, 1e-3)
    P = PR.scotty(B=1, N=N, n=4)
    P.x0 = xref[:1].copy()
    P.offsets = np.zeros(1, dtype=np.int32)
    u0 = np.array([uref[0][0], 0.0])
    P.U0 = np.tile(u0, (1, N, 1))
    s = altro_b200.make_solver(P)
    s.SetState(xref[:N + 1])
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scotty_mpc.json")))
    c_u = 0.5 * u0 @ (Rd * u0)
    x = xref[0].copy()
    n_iter_match, max_dx = 0, 0.0
    Nsim = 200
    for it in range(Nsim):
        status = s.Solve()
        assert status[0] == 0
        n_iter_match += int(s.GetIterations()[0] == gold["solve_iters"][it])
        if s.GetIterations()[0] != gold["solve_iters"][it]:
            print(f"MPC solve {it}: {s.GetIterations()[0]} iterations, golden {gold['solve_iters'][it]}")
        u_mpc = s.GetInputs()[0, 0]
        x = oracle.model_dynamics(oracle.MODEL_BICYCLE4, [2.7, 1.5], x, u_mpc, h)
        max_dx = max(max_dx, np.abs(x - np.array(gold["state_trajectory"][it + 1])).max())
        for k in range(N + 1):
            xk = xref[k + it + 1]
            q = -(Qd * xk)
            c = -(0.5 * q @ xk) + (c_u if k < N else 0.0)
            s.UpdateLinearCosts(q, None, c, k, k + 1)
        s.SetInitialState(x)
        s.ShiftTrajectory()
    s.close()
    assert n_iter_match >= Nsim - 2, f"iteration counts equal on {n_iter_match}/{Nsim} MPC solves"
    assert max_dx < 1e-6, f"closed-loop state error {max_dx}"


def test_advance_window_matches_oracle_rewindowing(oracle):
    """altro_b200_advance_window (on-device re-windowing of the tracking cost) == SetLQRCost on
    the shifted window in the oracle."""
    P = PR.scotty(B=48, N=30, n=4)
    s = altro_b200.make_solver(P)
    s.Solve()
    s.AdvanceWindow(3)
    s.ResetDuals()
    s.SetInput(P.U0)
    st = s.Solve()
    gpu = dict(X=s.GetStates(), U=s.GetInputs(), status=st, iters=s.GetIterations(),
               cost=s.GetFinalObjective())
    s.close()
    P2 = PR.scotty(B=48, N=30, n=4)
    P2.offsets = P.offsets + 3
    ref = oracle.solve_batch(P2)
    compare(gpu, ref)


def test_results_do_not_depend_on_batch_neighbours(oracle):
    """A problem's solution is independent of where it sits in the batch (no cross-lane state)."""
    P = PR.bicycle(B=70, N=60, n=5)
    full = altro_b200.solve_problem(P)
    part = altro_b200.solve_problem(P.subset(33, 70))
    assert np.array_equal(full["X"][33:], part["X"]) and np.array_equal(full["iters"][33:], part["iters"])


@pytest.mark.parametrize("make", [
    lambda: PR.bicycle(B=200, N=100, n=5),
    lambda: PR.scotty(B=128, N=30, n=4),
    lambda: PR.pendulum(B=100, N=50, goal_constraint=True),
    lambda: PR.double_integrator(N=10, variant="usoc", B=5),
    lambda: PR.chain(B=64, n=6, m=2, N=50, control_box=True),
])
def test_phase_pipeline_equals_persistent_kernel(make):
    """The production path (two kernels per iteration: staged sweeps, line-search rounds with the
    rollout / follower / speculating warps, fused post-search pass, alpha=0 scan instead of a
    re-rollout) and the single persistent kernel are two schedules of the same arithmetic: results
    must be bit-identical."""
    P = make()
    a = altro_b200.solve_problem(P, mode=0)
    b = altro_b200.solve_problem(P, mode=1)
    for key in ("X", "U", "Y", "status", "iters", "merit_evals", "cost", "stat", "feas"):
        assert np.array_equal(a[key], b[key]), key


def test_duals_of_stopped_problems_are_not_touched_again():
    """A constrained problem that stops while its group mates keep iterating gets exactly ONE dual
    update at the stop (solver.cpp:474-489), like in the reference and in the persistent twin: the
    final duals, penalties and the NEXT warm-started solve (duals carry over between MPC solves,
    quirk Q14) must be bit-identical between the two schedules on a batch whose problems stop at
    very different iterations."""
    P = PR.scotty(B=160, N=30, n=4)
    out = []
    for mode in (0, 1):
        s = altro_b200.make_solver(P)
        s.SetSolveMode(mode)
        s.SetMpcCostUpdate(0)
        s.Solve()
        it1 = s.GetIterations()
        z1, rho1 = s.GetField("z"), s.GetPenalty()
        s.MpcStep()
        s.Solve()
        out.append(dict(it1=it1, z1=z1, rho1=rho1, X=s.GetStates(), U=s.GetInputs(), it2=s.GetIterations(),
                        z2=s.GetField("z"), cost=s.GetFinalObjective()))
        s.close()
    a, b = out
    assert a["it1"].max() - a["it1"].min() >= 3          # heterogeneous stops inside the groups
    assert np.abs(a["z1"]).max() > 0
    for key in a:
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("make", [
    lambda: PR.bicycle(B=1000, N=100, n=5),
    lambda: PR.scotty(B=700, N=30, n=4),
])
def test_results_do_not_depend_on_schedule(make):
    """Speculation width (candidate steps rolled out per round) and the pipelined sub-batch split
    only change WHEN work is done, never the arithmetic: results must be bit-identical, run to
    run and schedule to schedule."""
    P = make()
    ref = None
    for nslots, nsplit, nstore in [(1, 1, 0), (4, 1, 3), (4, 3, 0), (8, 2, 2), (4, 3, 1), (8, 8, 7), (6, 4, 5)]:
        s = altro_b200.make_solver(P, nslots=nslots)
        s.SetPipelineSplit(nsplit)
        s.SetCandidateStore(nstore)
        s.Solve()
        out = dict(X=s.GetStates(), U=s.GetInputs(), iters=s.GetIterations(), evals=s.GetMeritEvals(),
                   cost=s.GetFinalObjective(), status=s.GetStatus())
        s.close()
        if ref is None:
            ref = out
            continue
        for key in ref:
            assert np.array_equal(out[key], ref[key]), (key, nslots, nsplit, nstore)


@pytest.mark.parametrize("make,model", [
    (lambda: PR.bicycle(B=40, N=20, n=5, iterations_max=3), PR.MODEL_BICYCLE5),
    (lambda: PR.scotty(B=40, N=15, n=4, iterations_max=3), PR.MODEL_BICYCLE4),
    (lambda: PR.pendulum(B=33, N=20, iterations_max=3), PR.MODEL_PENDULUM),
])
def test_knotpoint_views_match_oracle_model(oracle, make, model):
    """KnotPointData host views (altro_b200_get_field): A_, B_ of the accepted trajectory are the
    discrete Jacobians of the oracle's model at (x_k, u_k) -- also for the models whose Jacobian is
    stored packed and re-expanded on load -- and x_{k+1} = f(x_k, u_k)."""
    P = make()
    s = altro_b200.make_solver(P)
    s.Solve()
    X, U, A, B = s.GetField("x"), s.GetField("u"), s.GetField("A"), s.GetField("B")
    n, m = P.n, P.m
    assert A.shape == (P.B, P.N + 1, n * n) and B.shape == (P.B, P.N + 1, n * m)
    for b in (0, 7, P.B - 1):
        for k in (0, P.N // 2, P.N - 1):
            J = oracle.model_jacobian(model, P.model_params, X[b, k], U[b, k], P.h)  # n x (n+m)
            Ak = A[b, k].reshape(n, n, order="F")
            Bk = B[b, k].reshape(n, m, order="F")
            np.testing.assert_allclose(Ak, J[:, :n], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(Bk, J[:, n:], rtol=1e-12, atol=1e-14)
            xn = oracle.model_dynamics(model, P.model_params, X[b, k], U[b, k], P.h)
            np.testing.assert_allclose(X[b, k + 1], xn, rtol=1e-12, atol=1e-13)
    hist = s.GetLinesearchHistogram()
    assert hist.sum() > 0
    s.close()


def test_on_device_mpc_reproduces_the_reference_golden_run():
    """The reference's 200-step receding-horizon run (test/bicycle_test.cpp:247-337, golden
    test/scotty_mpc.json) driven ENTIRELY by altro_b200_mpc_step -- no host round trip between
    solves: the window moves with the reference's own cost update (UpdateLinearCosts(q, nullptr, c):
    r frozen, c_u frozen, :317-328), x0 <- x_[1] (the plant is the model, :312), ShiftTrajectory.
    A batch of replicas of the single reference problem; every replica must do what the golden
    file records."""
    import json, os
    xref, uref, h = PR.load_scotty()
    n, m, N, B, Nsim = 4, 2, 30, 40, 200
    Rd = np.full(m, 1e-3)
    P = PR.scotty(B=B, N=N, n=4)
    P.x0 = np.tile(xref[0], (B, 1))
    P.offsets = np.zeros(B, dtype=np.int32)
    u0 = np.array([uref[0][0], 0.0])
    P.U0 = np.tile(u0, (B, N, 1))
    s = altro_b200.make_solver(P)
    s.SetState(xref[:N + 1])
    s.SetMpcCostUpdate(1, 0.5 * u0 @ (Rd * u0))          # c_u, bicycle_test.cpp:296
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scotty_mpc.json")))
    giters = np.array(gold["solve_iters"])
    gx = np.array(gold["state_trajectory"])
    differ, max_dx = [], 0.0
    for it in range(Nsim):
        status = s.Solve()
        assert (status == 0).all(), f"MPC solve {it}: {(status != 0).sum()} replicas not Success"
        iters = s.GetIterations()
        X = s.GetStates()
        assert (iters == iters[0]).all() and np.array_equal(X, np.broadcast_to(X[0], X.shape)), \
            f"replicas diverged at MPC solve {it}"
        if iters[0] != giters[it]:
            differ.append((it, int(iters[0]), int(giters[it])))
        # closed-loop state after applying u_[0]: x_[1] of the solved trajectory
        max_dx = max(max_dx, np.abs(X[0, 1] - gx[it + 1]).max())
        s.MpcStep()
    s.close()
    # The CPU oracle reproduces 200/200 and 1e-11 (tests/test_oracle_solver.py); the device differs
    # from glibc in the last bit of sin/cos, which the unregularised iteration amplifies on a few
    # warm-started solves that sit on an Armijo knife edge -- report which, never more than two
    print(f"MPC solves with a different iteration count (step, gpu, golden): {differ}; "
          f"max closed-loop state error {max_dx:.3e}")
    assert len(differ) <= 2, differ
    assert max_dx < 1e-6, max_dx


def test_on_device_mpc_step_equals_host_driven_loop():
    """altro_b200_mpc_step in its re-windowing mode (x0 <- x_[1], ShiftTrajectory, window + 1 with q,
    r and c all following the window) against the same receding-horizon loop driven from the host
    through SetInitialState / ShiftTrajectory / AdvanceWindow: bit-identical over several
    warm-started solves of a scotty batch.  (The reference's own update -- r frozen -- is pinned to
    the golden run above.)"""
    P = PR.scotty(B=96, N=30, n=4)
    a = altro_b200.make_solver(P)
    b = altro_b200.make_solver(P)
    a.SetMpcCostUpdate(0)
    for step in range(6):
        sa, sb = a.Solve(), b.Solve()
        assert np.array_equal(sa, sb)
        Xa, Xb = a.GetStates(), b.GetStates()
        assert np.array_equal(Xa, Xb) and np.array_equal(a.GetIterations(), b.GetIterations())
        assert np.array_equal(a.GetInputs(), b.GetInputs())
        a.MpcStep()
        b.SetInitialState(Xb[:, 1].copy())
        b.ShiftTrajectory()
        b.AdvanceWindow(1)
    assert (a.GetStatus() == 0).mean() > 0.9
    a.close()
    b.close()
