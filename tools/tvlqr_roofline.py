#!/usr/bin/env python
"""Roofline line of the standalone batched TVLQR sweep (altro_b200_tvlqr_ws_*): algorithmic HBM bytes
per knot (all input rows read once, K d P p written once) x knots x problems / CUDA-event time of
the kernel, against the measured copy bandwidth.  Usage: python tools/tvlqr_roofline.py [n m N B diag]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import altro_b200  # noqa: E402
import bench  # noqa: E402

dp = C.POINTER(C.c_double)


def main():
    n, m, N, B, diag = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (6, 2, 200, 32768, 1))]
    L = altro_b200.load_library()
    vp = C.c_void_p
    L.altro_b200_tvlqr_ws_create.restype = vp
    L.altro_b200_tvlqr_ws_create.argtypes = [C.c_int] * 4 + [C.c_bool, C.c_int]
    L.altro_b200_tvlqr_ws_upload.argtypes = [vp] + [dp] * 8
    L.altro_b200_tvlqr_ws_time_backward.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_float)]
    L.altro_b200_tvlqr_ws_bytes_per_knot.restype = C.c_long
    L.altro_b200_tvlqr_ws_bytes_per_knot.argtypes = [vp]
    L.altro_b200_tvlqr_ws_destroy.argtypes = [vp]
    rng = np.random.default_rng(0)
    ptr = lambda a: a.ctypes.data_as(dp)
    A = np.ascontiguousarray(np.eye(n) + 0.05 * rng.normal(size=(B, N, n, n)))
    Bm = rng.normal(size=(B, N, m, n))
    f = 0.1 * rng.normal(size=(B, N, n))
    if diag:
        Q, R, H = 0.5 + rng.uniform(size=(B, N + 1, n)), 0.1 + rng.uniform(size=(B, N, m)), None
    else:
        Q = np.ascontiguousarray(np.broadcast_to(np.eye(n), (B, N + 1, n, n)))
        R = np.ascontiguousarray(np.broadcast_to(np.eye(m), (B, N, m, m)))
        H = np.zeros((B, N, n, m))
    q, r = rng.normal(size=(B, N + 1, n)), rng.normal(size=(B, N, m))
    w = vp(L.altro_b200_tvlqr_ws_create(B, n, m, N, bool(diag), 0))
    assert w.value, "shape not compiled in / no device"
    assert L.altro_b200_tvlqr_ws_upload(w, ptr(A), ptr(Bm), ptr(f), ptr(Q), ptr(R), ptr(H) if H is not None else None,
                                        ptr(q), ptr(r)) == 0
    ms = C.c_float()
    assert L.altro_b200_tvlqr_ws_time_backward(w, 0.0, 10, C.byref(ms)) == 0
    per_knot = L.altro_b200_tvlqr_ws_bytes_per_knot(w)
    L.altro_b200_tvlqr_ws_destroy(w)
    peak, src = bench.peak_hbm()
    byts = float(per_knot) * N * B
    gbs = byts / (ms.value * 1e-3) / 1e9
    print(json.dumps({"kernel": "k_tvlqr_backward_rec (standalone batched tvlqr_BackwardPass)", "n": n, "m": m, "N": N,
                      "batch": B, "is_diag": bool(diag), "ms_per_launch": ms.value, "bytes_per_knot": per_knot,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "peak_source": src}}))


if __name__ == "__main__":
    main()
