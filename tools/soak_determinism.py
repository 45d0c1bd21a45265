#!/usr/bin/env python
"""Full-size determinism soak: the bicycle batch (16 384 problems) and the scotty batch (8 192)
solved repeatedly under different schedules (sub-batch split 1 / 8 / 16, follower on / off) must
give bit-identical states, inputs, iteration counts and duals every time -- any ordering bug in the
staging pipelines of the forward kernel would show up as run-to-run differences."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import altro_b200  # noqa: E402
from altro_b200 import problems as PR  # noqa: E402


def run(P, split, follower):
    os.environ["ALTRO_B200_FOLLOWER"] = str(follower)
    s = altro_b200.make_solver(P)
    s.SetPipelineSplit(split)
    out = []
    for _ in range(3):
        s.ResetTrajectory()
        s.ResetDuals()
        s.Solve()
        out.append((s.GetStates().copy(), s.GetInputs().copy(), s.GetIterations().copy(), s.GetStatus().copy()))
    s.close()
    return out


def main():
    bad = 0
    for name, P in (("bicycle", PR.bicycle(B=16384, N=100, n=5)), ("scotty", PR.scotty(B=8192, N=50, n=5))):
        ref = None
        runs = 0
        for follower in (1, 0):
            for split in (1, 8, 16):
                for X, U, it, st in run(P, split, follower):
                    runs += 1
                    if ref is None:
                        ref = (X, U, it, st)
                        continue
                    same = all(np.array_equal(a, b) for a, b in zip(ref, (X, U, it, st)))
                    if not same:
                        bad += 1
                        print(f"{name}: follower={follower} split={split} differs from the first run "
                              f"(max |dX| {np.abs(X - ref[0]).max():.3e}, iterations equal: {np.array_equal(it, ref[2])})")
        print(f"{name}: {runs} solves of {P.B} problems under 6 schedules, mismatching runs: {bad}")
    print("determinism soak", "FAILED" if bad else "OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
