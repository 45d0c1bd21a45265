#!/bin/bash
mkdir -p gpurun_out
T=r02e
# (1) the GPU test suite
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread > gpurun_out/${T}_gputests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gputests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_gputests.log | tail -30
# (2) schedule experiments (resident step time, 3 solves each)
L=gpurun_out/${T}_sched.log
: > $L
for cfg in "8 6" "16 6" "32 6" "16 4" "16 8" "32 8" "8 3"; do
  set -- $cfg
  timeout 60 python tools/diag_hang.py 16384 $1 $2 >> $L 2>&1
done
cat $L
# (3) standalone TVLQR roofline
timeout 120 python tools/tvlqr_roofline.py 6 2 200 32768 1 > gpurun_out/${T}_tvlqr_6_2_200.json 2>&1
timeout 120 python tools/tvlqr_roofline.py 4 2 200 32768 1 > gpurun_out/${T}_tvlqr_4_2_200.json 2>&1
timeout 120 python tools/tvlqr_roofline.py 6 4 200 32768 0 > gpurun_out/${T}_tvlqr_6_4_200_dense.json 2>&1
cat gpurun_out/${T}_tvlqr_*.json
# (4) n = 12 with the team sweep
timeout 200 python tools/phase_profile.py chain12 4096 0 1 > gpurun_out/${T}_phase_chain12.json 2>&1
head -c 1500 gpurun_out/${T}_phase_chain12.json
# (5) ncu: launch list of the bench command + full captures
timeout 600 bash tools/ncu_capture.sh ${T} bicycle > gpurun_out/${T}_ncu.log 2>&1
tail -5 gpurun_out/${T}_ncu.log
