"""BASELINE.json's full-size configurations on the GPU, checked through properties that do not
need the CPU oracle to solve the whole batch (plus an oracle comparison on a sample):

  * the returned states ARE the rollout of the returned controls (re-simulating U from x0 with the
    device model reproduces X bit for bit);
  * problems reported Success satisfy the reference's convergence criteria
    (|stationarity| < tol_stationarity, feasibility < tol_primal_feasibility, solver.cpp:464);
  * iteration counts respect iterations_max (+1 when exhausted, quirk Q4) and statuses are valid;
  * inequality constraints of converged problems hold to the feasibility tolerance;
  * a re-solve warm-started at the solution stops (almost) at once (idempotence);
  * permuting the batch permutes the answers (no cross-talk between lanes / groups / sub-batches).
"""
import numpy as np
import pytest

import altro_b200
from altro_b200 import problems as PR
from parity_util import compare, compare_all

pytestmark = pytest.mark.gpu


def solve_full(P):
    s = altro_b200.make_solver(P)
    s.Solve()
    out = dict(X=s.GetStates(), U=s.GetInputs(), status=s.GetStatus(), iters=s.GetIterations(),
               evals=s.GetMeritEvals(), cost=s.GetFinalObjective(), stat=s.GetStationarity(),
               feas=s.GetPrimalFeasibility())
    return s, out


def check_common(P, s, out):
    imax = P.options.get("iterations_max", 200)
    st = out["status"]
    assert set(np.unique(st)) <= {0, 1, 2}
    assert np.all(out["iters"][st != 2] <= imax) and np.all(out["iters"][st == 2] == imax + 1)
    ok = st == 0
    assert ok.any()
    assert np.all(np.abs(out["stat"][ok]) < 1e-4) and np.all(out["feas"][ok] < 1e-4)
    assert np.all(np.isfinite(out["X"][ok])) and np.all(np.isfinite(out["cost"][ok]))
    # X is the rollout of U: OpenLoopRollout re-simulates the working inputs from x0
    s.OpenLoopRollout()
    X2 = s.GetStates()
    fin = np.isfinite(out["X"]).all(axis=(1, 2))
    assert np.array_equal(X2[fin], out["X"][fin])
    np.testing.assert_array_equal(out["X"][:, 0, :], P.x0)


def test_bicycle_16384():
    """BASELINE configs[2]: bicycle n=5 m=2 N=100, batch 16384 random goals."""
    P = PR.bicycle(B=16384, N=100, n=5)
    s, out = solve_full(P)
    check_common(P, s, out)
    ok = out["status"] == 0
    # idempotence: warm-started at the solution, a converged problem stops at once -- for the ~90 %
    # of them whose first unregularised Newton step from the solution is accepted; the CPU oracle
    # re-converges on the same 90 % (the rest wander off: reg = 0, solver.cpp:363)
    idx = np.nonzero(ok)[0][:2048]
    P2 = P.subset(0, P.B)
    P2.x0, P2.xref, P2.uref, P2.B = P.x0[idx], P.xref[idx], P.uref[idx], idx.size
    P2.U0, P2.U0_per_problem = out["U"][idx], True
    s2, out2 = solve_full(P2)
    again = out2["status"] == 0
    assert again.mean() > 0.85 and np.median(out2["iters"][again]) <= 2
    rel = np.abs(out2["cost"][again] - out["cost"][idx][again]) / np.abs(out["cost"][idx][again])
    assert np.median(rel) < 1e-6 and np.quantile(rel, 0.9) < 1e-3, (np.median(rel), np.quantile(rel, 0.9))
    s.close()
    s2.close()


def test_bicycle_permutation_invariance():
    P = PR.bicycle(B=4096, N=100, n=5)
    _, a = solve_full(P)
    perm = np.random.default_rng(5).permutation(P.B)
    Q = P.subset(0, P.B)
    Q.x0, Q.xref, Q.uref = P.x0[perm], P.xref[perm], P.uref[perm]
    _, b = solve_full(Q)
    for key in ("X", "U", "status", "iters", "evals", "cost"):
        assert np.array_equal(b[key], a[key][perm]), key


def test_bicycle_sample_against_oracle(oracle):
    """512 problems spread over the full batch (different groups and sub-batches) vs the oracle."""
    P = PR.bicycle(B=16384, N=100, n=5)
    _, out = solve_full(P)
    idx = np.arange(0, P.B, 32)
    Q = P.subset(0, P.B)
    Q.x0, Q.xref, Q.uref, Q.B = P.x0[idx], P.xref[idx], P.uref[idx], idx.size
    ref = oracle.solve_batch(Q)
    gpu = {k: out[k][idx] for k in ("X", "U", "status", "iters", "cost")}
    compare(gpu, ref, tail_frac=0.02)


def test_pendulum_4096(oracle):
    """BASELINE configs[1]: pendulum n=2 m=1 N=100, batch 4096 perturbed x0 -- whole batch vs oracle."""
    P = PR.pendulum(B=4096, N=100)
    s, out = solve_full(P)
    check_common(P, s, out)
    ref = oracle.solve_batch(P)
    compare({k: out[k] for k in ("X", "U", "status", "iters", "cost")}, ref)
    s.close()


def test_scotty_8192(oracle):
    """BASELINE configs[3], one GPU's share: scotty tracking MPC n=5 N=50 with the steering bound."""
    P = PR.scotty(B=8192, N=50, n=5)
    s, out = solve_full(P)
    check_common(P, s, out)
    ok = out["status"] == 0
    dmax = 60 * np.pi / 180
    assert np.all(np.abs(out["X"][ok][:, :, 3]) <= dmax + 1e-4)
    idx = np.arange(0, P.B, 16)
    Q = P.subset(0, P.B)
    Q.x0, Q.offsets, Q.U0, Q.B = P.x0[idx], P.offsets[idx], P.U0[idx], idx.size
    ref = oracle.solve_batch(Q)
    compare({k: out[k][idx] for k in ("X", "U", "status", "iters", "cost")}, ref, tail_frac=0.02)
    s.close()


@pytest.mark.parametrize("n,m,N", [(4, 2, 50), (6, 4, 200), (12, 4, 50)])
def test_chain_sweep_32768(oracle, n, m, N):
    """BASELINE configs[4] dimension sweep, batch 32768: properties + a 256-problem oracle sample."""
    P = PR.chain(B=32768, n=n, m=m, N=N)
    s, out = solve_full(P)
    check_common(P, s, out)
    idx = np.arange(0, P.B, 128)
    Q = P.subset(0, P.B)
    Q.x0, Q.xref, Q.uref, Q.B = P.x0[idx], P.xref[idx], P.uref[idx], idx.size
    ref = oracle.solve_batch(Q)
    compare({k: out[k][idx] for k in ("X", "U", "status", "iters", "cost")}, ref, tail_frac=0.02)
    s.close()


@pytest.mark.parametrize("name,make", [
    ("bicycle 16384", lambda: PR.bicycle(B=16384, N=100, n=5)),
    ("scotty 8192", lambda: PR.scotty(B=8192, N=50, n=5)),
])
def test_every_problem_matches_oracle_at_low_iteration_counts(oracle, name, make):
    """The WHOLE headline batch against the oracle, every problem regardless of its final status,
    with iterations_max = 1, 3, 10: before the unregularised iteration (reg = 0, solver.cpp:363) has
    had time to amplify last-bit differences of sin/cos, GPU and CPU must agree on the iteration
    count, the status and -- to 1e-9 -- on states, inputs and merit value.  A line-search decision
    sitting exactly on the Armijo threshold can still flip on a handful of the 16384 problems (the
    step length then differs by a factor two); those are counted, printed and bounded."""
    P = make()
    s = altro_b200.make_solver(P)
    growth = {}
    for itmax in (1, 3, 10):
        P.options = dict(P.options, iterations_max=itmax)
        s.SetOptions(altro_b200.default_options(**P.options))
        s.ResetTrajectory()
        s.ResetDuals()
        s.Solve()
        gpu = dict(X=s.GetStates(), U=s.GetInputs(), status=s.GetStatus(), iters=s.GetIterations(),
                   cost=s.GetFinalObjective())
        ref = oracle.solve_batch(P)
        rep = compare_all(gpu, ref)
        growth[itmax] = rep
        print(f"{name}, iterations_max={itmax}: {rep}")
        n = rep["n"]
        assert rep["finite_mismatch"] <= max(1, n // 2000)
        # measured on the B200: iterations_max = 1: every problem equal in status / iterations; 4 of
        # 16384 (bicycle) / 34 of 8192 (scotty) apart by more than 1e-9 (an Armijo decision of a deep
        # halving, where the decrease alpha * phi' is of the size of the rounding of phi, flipped:
        # alpha differs by a factor two); iterations_max = 3: 121 (0.7 %) on the bicycle batch;
        # median difference 1e-14
        if itmax <= 3:
            assert rep["same_status_and_iterations"] >= n - max(2, n // 2000), rep
            assert rep["n_above_1e9"] <= (n // 100 if itmax == 1 else n // 25), rep
            assert rep["median_err"] < 1e-12 and rep["p99_state_err"] < 1e-6, rep
        else:
            assert rep["same_status_and_iterations"] >= 0.97 * n, rep
            assert rep["n_above_1e6"] <= 0.10 * n and rep["median_err"] < 1e-9, rep
    s.close()
