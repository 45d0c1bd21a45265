#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02k}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_general.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests.log 2>&1
rc=$?; echo "parity+general rc=$rc" >> gpurun_out/${T}_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_tests.log | tail -6
ALTRO_B200_INLINE_DERIV=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "persistent or schedule or stopped or mpc" --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests_il1.log 2>&1
echo "inline=1 twin rc=$?" >> gpurun_out/${T}_tests_il1.log
tail -3 gpurun_out/${T}_tests_il1.log
L=gpurun_out/${T}_sched.log
: > $L
for fp in 0 1; do for il in 0 1; do
  echo "fused_post=$fp inline=$il" >> $L
  ALTRO_B200_FUSED_POST=$fp ALTRO_B200_INLINE_DERIV=$il timeout 60 python tools/diag_hang.py 16384 8 6 >> $L 2>&1
done; done
cat $L
for fp in 0 1; do
ALTRO_B200_FUSED_POST=$fp ALTRO_B200_INLINE_DERIV=1 timeout 120 python tools/phase_profile.py bicycle 16384 0 1 > gpurun_out/${T}_phase_bicycle_fp$fp.json 2>> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_phase_bicycle_fp$fp.json"))
print("fused_post=$fp", {k: round(v["ms"],2) for k,v in d["phases"].items()}, "wall", round(d["wall_ms_unprofiled"],2))
PY
done
for wl in scotty pendulum chain6; do for fp in 0 1; do
  ALTRO_B200_FUSED_POST=$fp timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_fp$fp.json 2>> gpurun_out/${T}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${wl}_fp$fp.json"))
    print("fused_post=$fp ${wl}", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("${wl} failed", e)
PY
done; done
tail -3 gpurun_out/${T}_bench.err
