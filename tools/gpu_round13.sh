#!/bin/bash
T=${1:-r02q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests.log 2>&1
rc=$?; echo "parity rc=$rc" >> gpurun_out/${T}_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_tests.log | tail -6
for sr in 1 0; do
echo "spec_round1=$sr"
ALTRO_B200_SPEC_ROUND1=$sr ALTRO_B200_PROF_DUMP=1 timeout 120 python tools/phase_profile.py bicycle 16384 0 1 > gpurun_out/${T}_phase_bicycle_sr$sr.json 2> gpurun_out/${T}_prof_sr$sr.txt
python - <<PY
import json,re
d=json.load(open("gpurun_out/${T}_phase_bicycle_sr$sr.json"))
print({k: round(v["ms"],2) for k,v in d["phases"].items()}, "wall", round(d["wall_ms_unprofiled"],2), "evals", d["mean_evals"])
tot=[0]*9
for l in open("gpurun_out/${T}_prof_sr$sr.txt"):
    m=re.findall(r"\d+", l)
    if l.startswith("fwd prof") and len(m)>=9:
        v=list(map(int,m))
        for i in range(9): tot[i]+=v[i]
if tot[8]: print("  wait-full share %.3f release/refill share %.3f pass cycles %d" % (tot[6]/tot[8], tot[7]/tot[8], tot[8]))
PY
ALTRO_B200_SPEC_ROUND1=$sr python tools/diag_hang.py 16384 8 6
ALTRO_B200_SPEC_ROUND1=$sr python tools/diag_hang.py 16384 8 8
for wl in scotty; do
  ALTRO_B200_SPEC_ROUND1=$sr timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_sr$sr.json 2>> gpurun_out/${T}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${wl}_sr$sr.json"))
    print("${wl}", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("${wl} failed", e)
PY
done
done
