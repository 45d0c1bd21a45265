"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/altro_b200.h declares, refuses to run without a device (no CPU fallback), and the host
logic that needs no GPU (index ranges, sizes) behaves like the reference."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "altro_b200.h")
LIB = os.path.join(ROOT, "altro_b200", "libaltro_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:altro_b200|tvlqr)_\w+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from altro_b200 import build
        build.build()
    return C.CDLL(LIB)


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/altro_b200.h but not exported: {missing}"


def test_exported_symbols_are_unmangled_c(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in declared_symbols():
        assert n in exported


def test_no_cpu_fallback(lib):
    import altro_b200
    if altro_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(altro_b200.AltroB200Error) as e:
        altro_b200.BatchSolver(10, 4)
    assert e.value.code == altro_b200.ErrorCodes.NoDevice
    # the batched tvlqr entry points refuse as well
    n, m, N, B = 2, 1, 3, 1
    z = lambda *s: np.zeros(s)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L = altro_b200.load_library()
    st = np.zeros(B, dtype=np.int32)
    rc = L.altro_b200_tvlqr_backward_batch(B, n, m, N, d(z(B, N, n * n)), d(z(B, N, n * m)), d(z(B, N, n)),
                                           d(z(B, N + 1, n)), d(z(B, N, m)), None, d(z(B, N + 1, n)),
                                           d(z(B, N, m)), 0.0, True, d(z(B, N, m * n)), d(z(B, N, m)),
                                           d(z(B, N + 1, n * n)), d(z(B, N + 1, n)), d(z(B, 2)),
                                           st.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == altro_b200.ErrorCodes.NoDevice


def test_tvlqr_total_mem_size_matches_reference_contract(lib, oracle):
    """tvlqr_TotalMemSize (tvlqr.cpp:18-63) needs no device: compare with the oracle restatement
    and with the hand-laid layout of tvlqr_test.cpp:67-72,167."""
    lib.tvlqr_TotalMemSize.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_bool]
    for n, m, N in [(4, 2, 10), (5, 2, 100), (12, 4, 7)]:
        nx = (C.c_int * (N + 1))(*([n] * (N + 1)))
        nu = (C.c_int * N)(*([m] * N))
        for diag in (True, False):
            assert lib.tvlqr_TotalMemSize(nx, nu, N, diag) == \
                oracle.lib().oracle_tvlqr_total_mem_size(nx, nu, N, int(diag))
    assert lib.tvlqr_TotalMemSize(None, None, 3, True) == 0


def test_error_strings_and_defaults(lib):
    import altro_b200
    L = altro_b200.load_library()
    assert b"no error" in L.altro_b200_error_string(0)
    assert b"CPU fallback" in L.altro_b200_error_string(100)
    o = altro_b200.default_options()
    # solver_options.hpp:18-33
    assert (o.iterations_max, o.tol_stationarity, o.tol_primal_feasibility) == (200, 1e-4, 1e-4)
    assert (o.penalty_initial, o.penalty_scaling, o.penalty_max) == (1.0, 10.0, 1e8)
    assert o.tol_meritfun_gradient == 1e-8 and o.use_backtracking_linesearch == 0


def test_facade_header_compiles_and_links(tmp_path):
    """A consumer translation unit written against the reference API compiles against
    include/altro/altro_solver.hpp and links with the library (no GPU needed to build)."""
    src = tmp_path / "consumer.cpp"
    src.write_text(r'''
#include "altro/altro_solver.hpp"
#include <cstdio>
using namespace altro;
int main() {
  ALTROSolver solver(10);                       // examples/cmake/*/main.cpp:5-10
  ErrorCodes err = solver.SetDimension(4, 2);
  std::printf("SetDimension -> %d (%s)\n", (int)err, ErrorCodeToString(err));
  b200::DeviceDynamics model(b200::DeviceDynamics::DoubleIntegrator, {2});
  err = solver.SetExplicitDynamics(model.Function(), model.Jacobian());
  std::printf("SetExplicitDynamics -> %d\n", (int)err);
  std::printf("N = %d, LastIndex = %d, AllIndices = %d\n", solver.GetHorizonLength(), LastIndex, AllIndices);
  return solver.GetHorizonLength() == 10 ? 0 : 1;
}
''')
    exe = tmp_path / "consumer"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", os.path.dirname(LIB), "-laltro_b200",
                        "-Wl,-rpath," + os.path.dirname(LIB)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "N = 10" in r.stdout
