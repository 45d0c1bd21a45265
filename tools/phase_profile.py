#!/usr/bin/env python
"""Per-phase timing of the solve pipeline (CUDA events around every launch; serialises the
pipeline, so the sum is an upper bound of the unprofiled step).  Usage:
  python tools/phase_profile.py [workload] [batch] [nslots|0] [nsplit]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import altro_b200  # noqa: E402
import bench  # noqa: E402



def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "bicycle"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    nslots = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) > 0 else None
    nsplit = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    P = bench.workload(wl, B, 0, 1)
    s = altro_b200.make_solver(P, nslots=nslots)
    s.SetPipelineSplit(nsplit)
    for _ in range(2):
        s.ResetTrajectory(); s.ResetDuals(); s.Solve()
    s.GetLinesearchHistogram(reset=True)
    s.SetProfiling(1)
    s.ResetTrajectory(); s.ResetDuals(); s.Solve()
    st, syncs = s.GetPhaseStats()
    tot = sum(v["ms"] for v in st.values())
    for v in st.values():
        v["share"] = v["ms"] / tot
        v["us_per_launch"] = 1e3 * v["ms"] / max(v["launches"], 1)
        v["ns_per_unit"] = 1e6 * v["ms"] / max(v["units"], 1)
    import time
    s.SetProfiling(0)
    walls = []
    for _ in range(3):
        s.ResetTrajectory(); s.ResetDuals(); s.Synchronize()
        t0 = time.perf_counter(); s.Solve(); walls.append(1e3 * (time.perf_counter() - t0))
    out = {"workload": wl, "B": B, "nsplit": nsplit, "wall_ms_unprofiled": min(walls), "total_ms": tot, "syncs": int(syncs),
           "mean_iters": float(s.GetIterations().mean()), "mean_evals": float(s.GetMeritEvals().mean()),
           "ls_hist": s.GetLinesearchHistogram().tolist(), "phases": st}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
