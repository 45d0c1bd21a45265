"""The device line-search state machine (altro_b200/csrc/linesearch.cuh, compiled for the host)
must take exactly the decisions of the reference's CubicLineSearch: same probe points, alpha,
status and evaluation count.  Checked against the oracle port and (when built) the reference's
own compiled linesearch.cpp, on the known answers of linesearch_tests.cpp and on randomised
merit functions."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from test_oracle_linesearch import _cubic, _quad, _random_merit

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def machine(oracle, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lsm") / "liblsm.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out,
                    os.path.join(HERE, "ls_machine_harness.cpp")], check=True)
    L = C.CDLL(out)
    O = oracle
    L.lsm_run.restype = C.c_double
    L.lsm_run.argtypes = [O.MERIT_CB, C.c_void_p] + [C.c_double] * 5 + [C.c_int] * 2 + \
        [O.iptr, O.iptr, O.dptr, O.dptr, O.iptr, O.iptr]

    def run(merit, alpha0, phi0, dphi0, c1=1e-4, c2=0.9, try_cubic_first=False, use_backtracking=False):
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
        a = L.lsm_run(O.MERIT_CB(cb), None, alpha0, phi0, dphi0, c1, c2, int(try_cubic_first),
                      int(use_backtracking), C.byref(st), C.byref(it), C.byref(ph), C.byref(dph),
                      C.byref(sd), C.byref(cv))
        return dict(alpha=a, status=st.value, iters=it.value, phi=ph.value, dphi=dph.value,
                    sufficient_decrease=bool(sd.value), curvature=bool(cv.value), alphas=alphas)

    return run


def test_machine_known_answers(machine):
    """linesearch_tests.cpp:134-270"""
    f = _quad(1.0, 1.1)
    r = machine(f, 1.0, *f(0.0, True), c1=1e-4, c2=0.01)
    assert r["iters"] == 3 and r["alpha"] == pytest.approx(1.1, rel=1e-15) and r["status"] == 1
    f = _quad(-1.0, -0.1)
    r = machine(f, 1.0, *f(0.0, True))
    assert r["alpha"] == 2.0 and r["status"] == 7 and r["sufficient_decrease"] and not r["curvature"]
    for c, c2, iters in [(1.2, 1e-3, 3), (1.8, 0.01, 4), (0.8, 0.01, 2), (0.01, 0.01, 2)]:
        f = _cubic(c)
        r = machine(f, 1.0, *f(0.0, True), c1=1e-4, c2=c2)
        assert r["iters"] == iters and abs(r["alpha"] - c) < 1e-6 and r["status"] == 1


def _same(a, b):
    assert a["alphas"] == b["alphas"]
    assert a["status"] == b["status"] and a["iters"] == b["iters"]
    assert (a["alpha"] == b["alpha"]) or (math.isnan(a["alpha"]) and math.isnan(b["alpha"]))
    if a["iters"] > 0:    # phi_ is uninitialised in the reference when nothing was evaluated
        assert a["phi"] == b["phi"] or (math.isnan(a["phi"]) and math.isnan(b["phi"]))
    assert a["sufficient_decrease"] == b["sufficient_decrease"] and a["curvature"] == b["curvature"]


def test_machine_matches_oracle_port_and_reference(oracle, machine):
    rng = np.random.default_rng(21)
    have_ref = oracle.ref_lib() is not None
    for trial in range(500):
        merit = _random_merit(rng)
        phi0, dphi0 = merit(0.0, True)
        if trial % 25 == 0:
            dphi0 = abs(dphi0)      # not a descent direction
        c2 = float(rng.choice([0.9, 0.1, 0.01, 1e-3]))
        for cubic_first in (False, True):
            for backtrack in (False, True):
                args = (merit, 1.0, phi0, dphi0, 1e-4, c2, cubic_first, backtrack)
                a = machine(*args)
                _same(a, oracle.linesearch_run(*args))
                if have_ref:
                    _same(a, oracle.ref_linesearch_run(*args))


def test_machine_hard_cases(oracle, machine):
    """Merit functions that force zoom exhaustion, tiny windows and backtracking exhaustion."""
    cases = [
        lambda x, w: (-x + 1e3 * x * x * (x > 1e-9), -1 + 2e3 * x * (x > 1e-9)),   # only tiny steps
        lambda x, w: (1.0 + x, -1.0),                                              # lying derivative
        lambda x, w: (math.cos(40 * x) - x, -40 * math.sin(40 * x) - 1),
        lambda x, w: (-1e-12 * x, -1e-12),
        lambda x, w: (float("nan"), float("nan")),
    ]
    for merit in cases:
        phi0, dphi0 = merit(0.0, True)
        if math.isnan(phi0):
            phi0, dphi0 = 1.0, -1.0
        for cubic_first in (False, True):
            for backtrack in (False, True):
                args = (merit, 1.0, phi0, dphi0, 1e-4, 0.9, cubic_first, backtrack)
                _same(machine(*args), oracle.linesearch_run(*args))
