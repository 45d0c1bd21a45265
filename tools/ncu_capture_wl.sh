#!/bin/bash
# ON THE GPU BOX: full ncu captures of the two phase kernels of another workload (mid-solve,
# sub-batch split off).  Usage: bash tools/ncu_capture_wl.sh TAG workload batch
T=${1:-r02}; WL=${2:-scotty}; B=${3:-8192}
mkdir -p gpurun_out
for K in k_phase_backward k_phase_forward; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 8 -c 2 -f \
      -o gpurun_out/${T}_${WL}_${K} python tools/phase_profile.py ${WL} ${B} 0 1 > gpurun_out/${T}_${WL}_${K}.log 2>&1
done
ls -la gpurun_out/${T}_${WL}_*.ncu-rep
