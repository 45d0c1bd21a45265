import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import altro_b200
from altro_b200 import problems as PR
def run(P, mode):
    s = altro_b200.make_solver(P, 0); s.SetSolveMode(mode); s.Solve()
    out = {k: s.GetField(k) for k in ("x","u","A","B","lx","lu","K","d","P","p")}
    s.close(); return out
P = PR.chain(B=40, n=12, m=4, N=50, control_box=True, iterations_max=1)
P.options["tol_meritfun_gradient"] = 1e30
a = run(P,0); b = run(P,1)
for k in a:
    d = np.abs(a[k]-b[k])
    if k in ("u","B","lu","K","d"): d = d[:, :-1]
    rows = np.nonzero(d.max(axis=(0,1))>1e-14)[0]
    knots = np.nonzero(d.max(axis=(0,2))>1e-14)[0]
    pbs = np.nonzero(d.max(axis=(1,2))>1e-14)[0]
    print(k, "maxdiff", d.max(), "rows:", rows[:12], len(rows), "knots:", knots[-6:], len(knots), "problems:", pbs[:10], len(pbs))
