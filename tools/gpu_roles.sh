#!/bin/bash
T=${1:-r02}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" ALTRO_B200_PROF_DUMP=1 timeout 120 python tools/phase_profile.py bicycle 16384 ${SLOTS:-0} 1 > gpurun_out/${T}_phase_$name.json 2> gpurun_out/${T}_prof_$name.txt
  python - <<PY
import json,re
d=json.load(open("gpurun_out/${T}_phase_$name.json"))
print("$name", {k: round(v["ms"],2) for k,v in d["phases"].items()}, "wall", round(d["wall_ms_unprofiled"],2))
tot=[0]*17
for l in open("gpurun_out/${T}_prof_$name.txt"):
    m=re.findall(r"\d+", l)
    if l.startswith("fwd prof") and len(m)>=17:
        v=list(map(int,m))
        for i in range(17): tot[i]+=v[i]
if tot[8]:
    print("   wait-full %.3f release %.3f" % (tot[6]/tot[8], tot[7]/tot[8]))
    for r in range(4):
        if tot[13+r]: print("   round %d: passes %d, cycles/knot %.0f" % (r, tot[13+r], tot[9+r]/tot[13+r]/100))
PY
}
run w0 ALTRO_B200_PROF_TID=0
run follower ALTRO_B200_PROF_TID=32
run spec1 ALTRO_B200_PROF_TID=64
run spec4 ALTRO_B200_PROF_TID=160
python tools/diag_hang.py 16384 8 6
