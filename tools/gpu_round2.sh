#!/bin/bash
# stage 1: does the big batch run at all (both Riccati schedules)?  abort early if not
mkdir -p gpurun_out
L=gpurun_out/r02d_diag.log
: > $L
run() { echo "== $* (team=$ALTRO_B200_BACKWARD_TEAM)" >> $L; timeout 60 python tools/diag_hang.py "$@" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L; return $rc; }
run 4768 1 6 || { tail -20 $L; exit 1; }
run 16384 1 6 || { tail -20 $L; exit 1; }
run 16384 4 6 || { tail -20 $L; exit 1; }
run 16384 8 6
ALTRO_B200_BACKWARD_TEAM=1 run 16384 4 6
ALTRO_B200_BACKWARD_TEAM=1 run 16384 8 6
ALTRO_B200_BACKWARD_TEAM=1 run 16384 1 6
tail -30 $L
# stage 2: the GPU test suite (default schedule), then the twin/schedule subset with the team sweep
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread > gpurun_out/r02d_gputests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r02d_gputests.log
grep -v "^  File\|^    " gpurun_out/r02d_gputests.log | tail -40
ALTRO_B200_BACKWARD_TEAM=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_general.py -q --timeout 200 --timeout-method=thread > gpurun_out/r02d_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/r02d_team_tests.log
grep -v "^  File\|^    " gpurun_out/r02d_team_tests.log | tail -15
# stage 3: phase profile + bench
for ns in 1 4; do
  timeout 200 python tools/phase_profile.py bicycle 16384 0 $ns > gpurun_out/r02d_phase_bicycle_split$ns.json 2>&1
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -c 1200 gpurun_out/r02d_bench.json
