#!/bin/bash
T=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests.log 2>&1
rc=$?; echo "parity rc=$rc" >> gpurun_out/${T}_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_tests.log | tail -6
ALTRO_B200_PROF_DUMP=1 timeout 120 python tools/phase_profile.py bicycle 16384 0 1 > gpurun_out/${T}_phase_bicycle.json 2> gpurun_out/${T}_prof.txt
python - <<PY
import json,re
d=json.load(open("gpurun_out/${T}_phase_bicycle.json"))
print({k: round(v["ms"],2) for k,v in d["phases"].items()}, "wall", round(d["wall_ms_unprofiled"],2), "evals", d["mean_evals"])
tot=[0]*9
for l in open("gpurun_out/${T}_prof.txt"):
    m=re.findall(r"\d+", l)
    if l.startswith("fwd prof") and len(m)>=9:
        v=list(map(int,m))
        for i in range(9): tot[i]+=v[i]
print("  sum: iters,ctas,ns(rollout,expand,dphi,criteria), cycles(full,writable,pass):", tot)
if tot[8]: print("  wait-full share %.3f release/refill share %.3f" % (tot[6]/tot[8], tot[7]/tot[8]))
PY
head -2 gpurun_out/${T}_prof.txt
for d in 4 8; do
ALTRO_B200_FWD_DEPTH=$d python tools/diag_hang.py 16384 8 6
done
for wl in scotty pendulum chain6; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}.json 2>> gpurun_out/${T}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${wl}.json"))
    print("${wl}", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("${wl} failed", e)
PY
done
