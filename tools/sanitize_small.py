"""Small solves of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import altro_b200  # noqa: E402
from altro_b200 import problems as PR  # noqa: E402

for P in (PR.bicycle(B=70, N=20, n=5, iterations_max=4), PR.scotty(B=40, N=12, n=5, iterations_max=4),
          PR.pendulum(B=33, N=15, iterations_max=3), PR.double_integrator(N=10, variant="usoc", B=3),
          PR.chain(B=40, n=12, m=4, N=10, control_box=True, iterations_max=2)):
    s = altro_b200.make_solver(P)
    s.SetPipelineSplit(2)
    s.Solve()
    print(P.name, s.GetIterations()[:4], s.GetStatus()[:4])
    s.close()
