#!/usr/bin/env python
"""BASELINE configs[4]: dimension sweep n in {4,6,12} x m in {2,4} x N in {50,200,500}, batch 32768,
pendulum-chain model, one GPU.  For every shape: solves/s with the batch resident in
HBM, ms per solve of the batch, mean iterations, and the dominant kernel with its achieved fraction
of the HBM copy peak (algorithmic bytes of bench.kernel_models / CUDA-event time of the kernel).
Hard instances (problems.chain(hard=True): 5-20 iterations).  Writes gpurun_out/r02_sweep_table.json
and prints a markdown table."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import altro_b200  # noqa: E402
import bench  # noqa: E402
from altro_b200 import problems as PR  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    peak, src = bench.peak_hbm()
    rows = []
    for n in (4, 6, 12):
        for m in (2, 4):
            for N in (50, 200, 500):
                P = PR.chain(B=B, n=n, m=m, N=N, hard=True)
                s = altro_b200.make_solver(P)
                s.Solve()                                   # warm-up
                ts = []
                for _ in range(2):
                    s.ResetTrajectory(); s.ResetDuals(); s.Synchronize()
                    t0 = time.perf_counter(); s.Solve(); ts.append(time.perf_counter() - t0)
                iters, evals, status = s.GetIterations(), s.GetMeritEvals(), s.GetStatus()
                s.SetPipelineSplit(1); s.SetProfiling(1)
                s.ResetTrajectory(); s.ResetDuals(); s.Solve()
                st, _ = s.GetPhaseStats()
                models = bench.kernel_models(P, iters, evals)
                ker = {}
                for ph in ("backward", "forward"):
                    if st[ph]["launches"]:
                        by = 8.0 * models[ph]["doubles"] * models[ph]["units"] + models[ph].get("extra_bytes", 0.0)
                        ker[ph] = dict(ms=st[ph]["ms"], gbs=by / (st[ph]["ms"] * 1e-3) / 1e9)
                dom = max(ker, key=lambda k: ker[k]["ms"])
                t = min(ts)
                rows.append(dict(n=n, m=m, N=N, B=B, solves_per_s=B / t, ms=1e3 * t, mean_iters=float(iters.mean()),
                                 success=float((status == 0).mean()), dominant=dom,
                                 backward_ms_per_launch=st["backward"]["ms"] / max(st["backward"]["launches"], 1),
                                 backward_gbs=ker["backward"]["gbs"], backward_frac=ker["backward"]["gbs"] / peak,
                                 dominant_share=ker[dom]["ms"] / sum(v["ms"] for v in ker.values()),
                                 dominant_gbs=ker[dom]["gbs"], dominant_frac=ker[dom]["gbs"] / peak,
                                 hbm_gb=s.DeviceBytes() / 1e9))
                s.close()
                print(json.dumps(rows[-1]), flush=True)
    out = dict(peak_gbs=peak, peak_source=src, rows=rows)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_sweep_table.json"), "w"), indent=1)
    print("| n | m | N | solves/s | ms/batch | iters | dominant kernel | share | GB/s | frac of peak | HBM GB |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['n']} | {r['m']} | {r['N']} | {r['solves_per_s']:.0f} | {r['ms']:.1f} | {r['mean_iters']:.2f} | "
              f"{r['dominant']} | {r['dominant_share']:.2f} | {r['dominant_gbs']:.0f} | {r['dominant_frac']:.3f} | {r['hbm_gb']:.1f} |")


if __name__ == "__main__":
    main()
