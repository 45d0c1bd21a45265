"""How sensitive is the reference algorithm to last-bit arithmetic changes?

The same oracle sources compiled without and with floating-point contraction (-std=c11 default off vs -ffp-contract=fast)
(liboracle.so vs liboracle_fma.so) are run on the bicycle / scotty workloads.  On most problems
they agree to ~1e-12, but a small fraction of the converged solves ends with a different iteration
count or status: the un-regularised AL-iLQR iteration is chaotic on hard instances.  This bounds
what any independent implementation (such as the CUDA path) can be asked to reproduce, and is the
justification of `tail_frac` in tests/parity_util.py.
"""
import numpy as np

from altro_b200 import problems as PR
from parity_util import STATE_TOL, errors


def _solve(oracle, variant, P):
    oracle.use_variant(variant)
    try:
        return oracle.solve_batch(P)
    finally:
        oracle.use_variant("liboracle.so")


def test_contraction_changes_a_small_tail(oracle):
    rows = []
    for P in (PR.bicycle(B=1536, N=100, n=5), PR.scotty(B=1536, N=50, n=5)):
        a = _solve(oracle, "liboracle.so", P)
        b = _solve(oracle, "liboracle_fma.so", P)
        conv = a["status"] == 0
        same = (a["iters"] == b["iters"]) & (a["status"] == b["status"])
        ex, eu, ec = errors(b, a)
        tail = conv & ~(same & (ex <= STATE_TOL))
        rows.append((P.name, int(conv.sum()), int(tail.sum()), float(np.median(ex[conv & same]))))
        # the bulk agrees tightly ...
        assert np.median(ex[conv & same]) < 1e-10
        # ... and the tail is small but is allowed to exist
        assert tail.sum() <= 0.02 * conv.sum()
    print("oracle self-consistency (no-fma vs fma):", rows)


def test_well_conditioned_problems_are_insensitive(oracle):
    for P in (PR.pendulum(B=256, N=100), PR.chain(B=64, n=6, m=2, N=50)):
        a = _solve(oracle, "liboracle.so", P)
        b = _solve(oracle, "liboracle_fma.so", P)
        assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["status"], b["status"])
        ex, eu, ec = errors(b, a)
        assert ex.max() < 1e-12 and ec.max() < 1e-12
