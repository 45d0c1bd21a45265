#!/bin/bash
# ON THE GPU BOX: the whole validation of a build -- every GPU test, the bench lines, the ncu
# launch list and full captures of the two phase kernels.  Usage: bash tools/gpu_full.sh TAG
T=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/${T}_gputests.log 2>&1
rc=$?; echo "gpu tests rc=$rc" >> gpurun_out/${T}_gputests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_gputests.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for wl in scotty pendulum chain6 chain12 scotty_mpc; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}.json 2>> gpurun_out/${T}_bench.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python - <<PY
import json
for wl in ("", "_scotty", "_pendulum", "_chain6", "_chain12", "_scotty_mpc", "_reference"):
    try:
        d=json.load(open("gpurun_out/${T}_bench%s.json" % wl))
        print(wl or "bicycle", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms  e2e", round(d["e2e"]["value"]),
              "roofline", round(d.get("roofline",{}).get("frac",0),3), "step", round(d.get("step_roofline",{}).get("frac",0),3))
    except Exception as e:
        print(wl, "failed", e)
PY
tail -3 gpurun_out/${T}_bench.err
bash tools/ncu_capture.sh ${T} bicycle > gpurun_out/${T}_ncu.log 2>&1
tail -4 gpurun_out/${T}_ncu.log
python __graft_entry__.py --smoke 2>&1 | tail -4
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_clocks.txt
