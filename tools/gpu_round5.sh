#!/bin/bash
mkdir -p gpurun_out
T=r02g
ALTRO_B200_BACKWARD_TEAM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_team_tests.log 2>&1
rc=$?; echo "team parity rc=$rc" >> gpurun_out/${T}_team_tests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_team_tests.log | tail -12
[ $rc -ne 0 ] && exit 1
L=gpurun_out/${T}_sched.log
: > $L
ALTRO_B200_BACKWARD_TEAM=1 timeout 60 python tools/diag_hang.py 16384 8 6 >> $L 2>&1
ALTRO_B200_BACKWARD_TEAM=0 timeout 60 python tools/diag_hang.py 16384 8 6 >> $L 2>&1
cat $L
timeout 200 python tools/phase_profile.py chain12 4096 0 1 > gpurun_out/${T}_phase_chain12.json 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_phase_chain12.json'))
print('chain12 wall', d['wall_ms_unprofiled'], {k: round(v['us_per_launch'],1) for k,v in d['phases'].items()})
PY
timeout 900 python tools/sweep_table.py 32768 > gpurun_out/${T}_sweep.log 2>&1
tail -22 gpurun_out/${T}_sweep.log
