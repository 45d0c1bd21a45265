#!/bin/bash
mkdir -p gpurun_out
T=r02i
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_general.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests.log 2>&1
rc=$?; echo "parity+general rc=$rc" >> gpurun_out/${T}_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_tests.log | tail -12
[ $rc -ne 0 ] && exit 1
L=gpurun_out/${T}_sched.log
: > $L
for il in 1 0; do
  ALTRO_B200_INLINE_DERIV=$il timeout 60 python tools/diag_hang.py 16384 8 6 >> $L 2>&1
  ALTRO_B200_INLINE_DERIV=$il timeout 60 python tools/diag_hang.py 16384 1 6 >> $L 2>&1
done
cat $L
for il in 1 0; do for wl in scotty pendulum chain6 scotty_mpc; do
  ALTRO_B200_INLINE_DERIV=$il timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_il$il.json 2>> gpurun_out/${T}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${wl}_il$il.json"))
    print("inline=$il ${wl}", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("${wl} failed", e)
PY
done; done
tail -3 gpurun_out/${T}_bench.err
timeout 200 python tools/phase_profile.py bicycle 16384 0 1 > gpurun_out/${T}_phase_bicycle_split1.json 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -4
