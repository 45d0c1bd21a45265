"""Oracle line search / cubic spline vs the reference's known answers
(src/linesearch/test/linesearch_tests.cpp) and vs the reference's own compiled sources
(oracle/_ref, built by oracle/Makefile `ref` when /root/reference is present)."""
import math

import numpy as np
import pytest

CS_NOERROR, CS_FOUND_MINIMUM, CS_INVALIDPOINTER, CS_SADDLEPOINT, CS_NOMINIMUM, \
    CS_IS_POSITIVE_QUADRATIC, CS_IS_LINEAR, CS_IS_CONSTANT, CS_UNEXPECTED_ERROR, CS_SAME_POINT = range(10)
MINIMUM_FOUND, HIT_MAX_STEPSIZE = 1, 7


def test_cubic_spline_known_answers(oracle):
    O = oracle
    # linesearch_tests.cpp:26-41 constant
    x, e1, e2, co = O.cubic_argmin(0, 1.2, 0, 1, 1.2, 0)
    assert e1 == CS_NOERROR and e2 == CS_IS_CONSTANT and math.isnan(x)
    assert co[1:] == (1.2, 0.0, 0.0, 0.0)
    # :43-58 linear
    x, e1, e2, co = O.cubic_argmin(0, 0, 1, 1, 1, 1)
    assert e2 == CS_IS_LINEAR and math.isnan(x) and co[1:] == (0.0, 1.0, 0.0, 0.0)
    # :60-77 positive quadratic
    x, e1, e2, co = O.cubic_argmin(0.3, 0, -1.0, 0.7, 0, 1.0)
    assert e2 == CS_FOUND_MINIMUM and co[4] == 0 and x == pytest.approx(0.5, abs=1e-15)
    # :79-94 negative quadratic
    x, e1, e2, co = O.cubic_argmin(0.3, 0, 1.0, 0.7, 0, -1.0)
    assert e2 == CS_IS_POSITIVE_QUADRATIC and math.isnan(x)
    # :96-107 cubic
    x, e1, e2, co = O.cubic_argmin(0, 0, -1, 1, 0, 2)
    assert co[4] == 1.0 and e2 == CS_FOUND_MINIMUM and abs(x - 0.5773502691896257) < 1e-10
    # :109-119 no minimum
    x, e1, e2, co = O.cubic_argmin(0, 0, -1, 1, -3, -10)
    assert e2 == CS_NOMINIMUM and math.isnan(x)
    # same point
    x, e1, e2, co = O.cubic_argmin(0.5, 0, -1, 0.5 + 1e-7, -3, -10)
    assert e1 == CS_SAME_POINT


def _quad(a, c):
    return lambda x, want: (a * (x - c) ** 2, 2 * a * (x - c))


def _cubic(c):
    return lambda x, want: ((x - c) ** 2 - (x - c) ** 3, 2 * (x - c) - 3 * (x - c) ** 2)


def test_linesearch_quadratic_known_answers(oracle):
    """linesearch_tests.cpp:134-210"""
    O = oracle
    f = _quad(1.0, 1.0)
    r = O.linesearch_run(f, 1.0, *f(0.0, True))
    assert r["iters"] == 1 and r["alpha"] == 1.0 and r["status"] == MINIMUM_FOUND
    assert r["sufficient_decrease"] and r["curvature"]
    f = _quad(1.0, 1.1)
    r = O.linesearch_run(f, 1.0, *f(0.0, True))
    assert r["iters"] == 1 and r["alpha"] == 1.0 and r["status"] == MINIMUM_FOUND
    r = O.linesearch_run(f, 1.0, *f(0.0, True), c1=1e-4, c2=0.01)
    assert r["iters"] == 3 and r["alpha"] == pytest.approx(1.1, rel=1e-15)
    assert r["status"] == MINIMUM_FOUND
    f = _quad(1.0, 0.8)
    r = O.linesearch_run(f, 1.0, *f(0.0, True), c1=1e-4, c2=0.1)
    assert r["alpha"] == pytest.approx(0.8, rel=1e-15) and r["status"] == MINIMUM_FOUND
    f = _quad(-1.0, -0.1)
    r = O.linesearch_run(f, 1.0, *f(0.0, True), c1=1e-4, c2=0.9)
    assert r["sufficient_decrease"] and not r["curvature"] and r["alpha"] == 2.0
    assert r["status"] == HIT_MAX_STEPSIZE


def test_linesearch_cubic_known_answers(oracle):
    """linesearch_tests.cpp:212-270: exact alpha and iteration counts 1,3,4,2,2"""
    O = oracle
    f = _cubic(1.0)
    r = O.linesearch_run(f, 1.0, *f(0.0, True))
    assert r["iters"] == 1 and r["alpha"] == 1.0 and r["status"] == MINIMUM_FOUND
    for c, c2, iters, tol in [(1.2, 1e-3, 3, 1e-15), (1.8, 0.01, 4, 1e-15), (0.8, 0.01, 2, 1e-15),
                              (0.01, 0.01, 2, None)]:
        f = _cubic(c)
        r = O.linesearch_run(f, 1.0, *f(0.0, True), c1=1e-4, c2=c2)
        assert r["iters"] == iters, (c, r)
        if tol:
            assert r["alpha"] == pytest.approx(c, rel=tol)
        else:
            assert abs(r["alpha"] - c) < 1e-6
        assert r["status"] == MINIMUM_FOUND and r["sufficient_decrease"] and r["curvature"]


def test_not_descent_direction(oracle):
    r = oracle.linesearch_run(_quad(1.0, -1.0), 1.0, 1.0, 2.0)
    assert r["alpha"] == 0.0 and r["status"] == 3 and r["iters"] == 0 and r["alphas"] == []


def _random_merit(rng):
    """Smooth 1-D functions with a descent direction at 0, varied shapes."""
    kind = rng.integers(0, 4)
    a, b, c, d = rng.uniform(0.2, 3), rng.uniform(-2, 2), rng.uniform(0.05, 2.5), rng.uniform(0.1, 4)
    if kind == 0:
        f = lambda x: a * (x - c) ** 2 + b
        g = lambda x: 2 * a * (x - c)
    elif kind == 1:
        f = lambda x: (x - c) ** 2 - 0.3 * (x - c) ** 3
        g = lambda x: 2 * (x - c) - 0.9 * (x - c) ** 2
    elif kind == 2:
        f = lambda x: -math.sin(d * x) * a + 0.1 * x * x
        g = lambda x: -d * math.cos(d * x) * a + 0.2 * x
    else:
        f = lambda x: a * math.exp(-d * x) + c * x * x + b * 0.01 * x
        g = lambda x: -a * d * math.exp(-d * x) + 2 * c * x + b * 0.01
    return lambda x, want: (f(x), g(x))


def test_port_matches_reference_build(oracle):
    """Randomised cross-check of oracle/linesearch_port.c against the reference's compiled
    linesearch.cpp + cubicspline.c: same alpha (bitwise), status, iteration count, probe points."""
    O = oracle
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    rng = np.random.default_rng(7)
    n_checked = 0
    for trial in range(400):
        merit = _random_merit(rng)
        phi0, dphi0 = merit(0.0, True)
        c2 = float(rng.choice([0.9, 0.1, 0.01]))
        for cubic_first in (False, True):
            for backtrack in (False, True):
                a = O.linesearch_run(merit, 1.0, phi0, dphi0, 1e-4, c2, cubic_first, backtrack)
                b = O.ref_linesearch_run(merit, 1.0, phi0, dphi0, 1e-4, c2, cubic_first, backtrack)
                assert a["alphas"] == b["alphas"]
                assert a["status"] == b["status"] and a["iters"] == b["iters"]
                assert (a["alpha"] == b["alpha"]) or (math.isnan(a["alpha"]) and math.isnan(b["alpha"]))
                assert a["phi"] == b["phi"]
                assert a["sufficient_decrease"] == b["sufficient_decrease"]
                assert a["curvature"] == b["curvature"]
                n_checked += 1
    assert n_checked == 1600


def test_port_cubic_matches_reference_build(oracle):
    O = oracle
    R = O.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    import ctypes as C
    rng = np.random.default_rng(11)
    for _ in range(2000):
        v = rng.uniform(-3, 3, 6)
        if rng.random() < 0.1:
            v[3] = v[0] + rng.uniform(-2e-6, 2e-6)
        e1, e2 = C.c_int(), C.c_int()
        co = (C.c_double * 5)()
        xr = R.ref_cubic_argmin(*v, C.byref(e1), C.byref(e2), co)
        xo, o1, o2, oc = O.cubic_argmin(*v)
        assert o1 == e1.value and o2 == e2.value
        if e1.value == 0:
            assert tuple(co) == oc
            assert (xo == xr) or (math.isnan(xo) and math.isnan(xr))
