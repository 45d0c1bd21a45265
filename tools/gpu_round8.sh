#!/bin/bash
mkdir -p gpurun_out
T=r02j
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_general.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests.log 2>&1
rc=$?; echo "parity+general rc=$rc" >> gpurun_out/${T}_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_tests.log | tail -6
[ $rc -ne 0 ] && exit 1
ALTRO_B200_INLINE_DERIV=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "persistent or schedule or stopped or mpc" --timeout 300 --timeout-method=thread > gpurun_out/${T}_tests_il1.log 2>&1
echo "inline=1 twin rc=$?" >> gpurun_out/${T}_tests_il1.log
tail -3 gpurun_out/${T}_tests_il1.log
L=gpurun_out/${T}_sched.log
: > $L
for il in 0 1; do
  ALTRO_B200_INLINE_DERIV=$il timeout 60 python tools/diag_hang.py 16384 8 6 >> $L 2>&1
done
cat $L
for wl in scotty pendulum chain6 scotty_mpc; do for il in -1 0 1; do
  if [ $il -ge 0 ]; then export ALTRO_B200_INLINE_DERIV=$il; else unset ALTRO_B200_INLINE_DERIV; fi
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_il$il.json 2>> gpurun_out/${T}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${wl}_il$il.json"))
    print("inline=$il ${wl}", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("${wl} failed", e)
PY
done; done
unset ALTRO_B200_INLINE_DERIV
tail -3 gpurun_out/${T}_bench.err
