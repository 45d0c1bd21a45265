#!/bin/bash
T=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_general.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests.log 2>&1
rc=$?; echo "parity+general rc=$rc" >> gpurun_out/${T}_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_tests.log | tail -6
ALTRO_B200_INLINE_DERIV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "persistent or schedule or stopped or mpc" --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests_il0.log 2>&1
echo "inline=0 twin rc=$?" >> gpurun_out/${T}_tests_il0.log
tail -3 gpurun_out/${T}_tests_il0.log
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
run follow ALTRO_B200_INLINE_DERIV=1
run lean ALTRO_B200_INLINE_DERIV=0
SLOTS=1 run lean_nospec ALTRO_B200_INLINE_DERIV=0
python tools/diag_hang.py 16384 8 6
ALTRO_B200_INLINE_DERIV=0 python tools/diag_hang.py 16384 8 6
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
