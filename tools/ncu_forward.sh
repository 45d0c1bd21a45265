#!/bin/bash
# ON THE GPU BOX: one ncu --set full capture (with source) of k_phase_forward mid-solve
TAG=${1:-r02n}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_phase_forward -s 6 -c 1 -f \
    -o gpurun_out/${TAG}_bicycle_k_phase_forward python tools/phase_profile.py bicycle 16384 0 1 > gpurun_out/${TAG}_ncu_fwd.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_fwd.log
ls -la gpurun_out/${TAG}*
