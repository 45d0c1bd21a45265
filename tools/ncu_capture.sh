#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): launch list of the bench command + ncu --set full captures of
# the phase kernels.  Outputs into gpurun_out/.
set -u
TAG=${1:-r01}
WL=${2:-bicycle}
mkdir -p gpurun_out
# (1) every launch of the bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload ${WL} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_${WL}.log 2>&1
# (2) full captures of the phase kernels (3 launches each, mid-solve, sub-batch pipelining off)
for K in k_phase_backward k_phase_forward; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 8 -c 2 -f \
      -o gpurun_out/${TAG}_${WL}_${K} python tools/phase_profile.py ${WL} 16384 0 1 > gpurun_out/${TAG}_${WL}_${K}.log 2>&1
done
ls -la gpurun_out | tail -20
