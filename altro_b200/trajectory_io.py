"""Trajectory files with the reference's keys (test/scotty.json, test/scotty_mpc.json) for one problem
or a batch, through the C ABI (include/altro_b200.h, section D; altro_b200/csrc/json_io.cpp).
Host-only: works without a CUDA device."""
import ctypes as C

import numpy as np

from .solver import AltroB200Error, dptr, iptr, load_library


def _lib():
    L = load_library()
    if not getattr(L, "_traj_bound", False):
        L.altro_b200_traj_open.restype = C.c_void_p
        L.altro_b200_traj_open.argtypes = [C.c_char_p, iptr]
        L.altro_b200_traj_dims.argtypes = [C.c_void_p, iptr, iptr, C.POINTER(C.c_float), iptr, iptr, iptr, iptr, iptr]
        L.altro_b200_traj_read.argtypes = [C.c_void_p, dptr, dptr, iptr, dptr]
        L.altro_b200_traj_close.argtypes = [C.c_void_p]
        L.altro_b200_traj_write.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, dptr,
                                            C.c_int, C.c_int, dptr, C.c_int, iptr, dptr]
        L._traj_bound = True
    return L


def read_trajectory(path):
    """-> dict(batch, N, tf, state_trajectory [B,knots,n], input_trajectory [B,knots,m],
    solve_iters [B,steps] | None, tracking_error [B,steps] | None); B = 1 for the reference's files."""
    L = _lib()
    err = C.c_int()
    f = L.altro_b200_traj_open(str(path).encode(), C.byref(err))
    if not f:
        raise AltroB200Error(err.value, f"(altro_b200_traj_open {path})")
    f = C.c_void_p(f)
    try:
        v = [C.c_int() for _ in range(7)]
        tf = C.c_float()
        L.altro_b200_traj_dims(f, C.byref(v[0]), C.byref(v[1]), C.byref(tf), C.byref(v[2]), C.byref(v[3]),
                               C.byref(v[4]), C.byref(v[5]), C.byref(v[6]))
        B, N, kx, n, ku, m, steps = [x.value for x in v]
        X = np.zeros((B, kx, n))
        U = np.zeros((B, ku, m))
        it = np.zeros((B, steps), dtype=np.int32)
        te = np.zeros((B, steps))
        L.altro_b200_traj_read(f, X.ctypes.data_as(dptr), U.ctypes.data_as(dptr), it.ctypes.data_as(iptr),
                               te.ctypes.data_as(dptr))
    finally:
        L.altro_b200_traj_close(f)
    return dict(batch=B, N=N, tf=float(tf.value), state_trajectory=X, input_trajectory=U,
                solve_iters=it if steps else None, tracking_error=te if steps else None)


def write_trajectory(path, N, tf, X, U, solve_iters=None, tracking_error=None):
    """X [knots,n] / U [knots,m] write the reference's single-problem layout; X [B,knots,n] /
    U [B,knots,m] a batch file ("batch": B)."""
    L = _lib()
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    batch = X.shape[0] if X.ndim == 3 else 0
    kx, n = X.shape[-2:]
    ku, m = U.shape[-2:]
    it = te = None
    steps = 0
    if solve_iters is not None:
        it = np.ascontiguousarray(solve_iters, dtype=np.int32)
        steps = it.shape[-1]
    if tracking_error is not None:
        te = np.ascontiguousarray(tracking_error, dtype=np.float64)
        steps = te.shape[-1]
    e = L.altro_b200_traj_write(str(path).encode(), batch, int(N), C.c_float(tf), kx, n, X.ctypes.data_as(dptr),
                                ku, m, U.ctypes.data_as(dptr), steps,
                                it.ctypes.data_as(iptr) if it is not None else None,
                                te.ctypes.data_as(dptr) if te is not None else None)
    if e:
        raise AltroB200Error(e, f"(altro_b200_traj_write {path})")
