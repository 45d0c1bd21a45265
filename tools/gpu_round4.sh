#!/bin/bash
mkdir -p gpurun_out
T=r02f
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x -k "persistent or schedule or stopped or mpc" --timeout 200 --timeout-method=thread > gpurun_out/${T}_twin.log 2>&1
rc=$?; echo "twin rc=$rc" >> gpurun_out/${T}_twin.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_twin.log | tail -12
[ $rc -ne 0 ] && exit 1
L=gpurun_out/${T}_sched.log
: > $L
for d in 4 8; do for st in 5 2 0; do
  ALTRO_B200_FWD_DEPTH=$d timeout 60 python tools/diag_hang.py 16384 8 6 0 30 $st >> $L 2>&1
done; done
ALTRO_B200_FWD_DEPTH=8 timeout 60 python tools/diag_hang.py 16384 1 6 0 30 5 >> $L 2>&1
ALTRO_B200_FWD_DEPTH=8 timeout 60 python tools/diag_hang.py 16384 4 6 0 30 5 >> $L 2>&1
ALTRO_B200_FWD_DEPTH=6 timeout 60 python tools/diag_hang.py 16384 8 6 0 30 5 >> $L 2>&1
cat $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread > gpurun_out/${T}_gputests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gputests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_gputests.log | tail -12
timeout 200 python tools/phase_profile.py bicycle 16384 0 1 > gpurun_out/${T}_phase_bicycle_split1.json 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.json
