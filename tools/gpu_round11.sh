#!/bin/bash
T=${1:-r02o}
mkdir -p gpurun_out
for sl in 0 2; do
ALTRO_B200_PROF_DUMP=1 timeout 120 python tools/phase_profile.py bicycle 16384 $sl 1 > gpurun_out/${T}_phase_bicycle_sl$sl.json 2> gpurun_out/${T}_prof_sl$sl.txt
python - <<PY
import json,re
d=json.load(open("gpurun_out/${T}_phase_bicycle_sl$sl.json"))
print("slots=$sl", {k: round(v["ms"],2) for k,v in d["phases"].items()}, "wall", round(d["wall_ms_unprofiled"],2), "evals", d["mean_evals"])
tot=[0]*9
for l in open("gpurun_out/${T}_prof_sl$sl.txt"):
    m=re.findall(r"\d+", l)
    if l.startswith("fwd prof") and len(m)>=9:
        v=list(map(int,m))
        for i in range(9): tot[i]+=v[i]
print("  sum: iters,ctas,ns(rollout,expand,dphi,criteria), cycles(full,writable,pass):", tot)
if tot[8]: print("  wait-full share %.3f wait-writable share %.3f" % (tot[6]/tot[8], tot[7]/tot[8]))
PY
head -3 gpurun_out/${T}_prof_sl$sl.txt
done
python tools/diag_hang.py 16384 8 6
