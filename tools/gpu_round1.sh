#!/bin/bash
# first GPU validation of the fused forward kernel: twin equality first (fast fail), full GPU suite, phase profile, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent or schedule" --timeout 240 --timeout-method=thread > gpurun_out/r02a_twin.log 2>&1
echo "twin rc=$?" >> gpurun_out/r02a_twin.log
tail -5 gpurun_out/r02a_twin.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread > gpurun_out/r02a_gputests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r02a_gputests.log
tail -15 gpurun_out/r02a_gputests.log
for ns in 1 2 4 8; do
  timeout 300 python tools/phase_profile.py bicycle 16384 0 $ns > gpurun_out/r02a_phase_bicycle_split$ns.json 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 1500 gpurun_out/r02a_bench.json
