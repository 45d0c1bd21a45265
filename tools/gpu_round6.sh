#!/bin/bash
mkdir -p gpurun_out
T=r02h
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread > gpurun_out/${T}_gputests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/${T}_gputests.log
grep -v "^  File\|^    \|^E    " gpurun_out/${T}_gputests.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
for wl in scotty scotty_mpc pendulum chain6; do
  timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 --cpu-seconds 6 > gpurun_out/${T}_bench_${wl}.json 2>> gpurun_out/${T}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${wl}.json"))
    print("${wl}", round(d["value"]), "solves/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]), "cpu", round(d.get("cpu_baseline",{}).get("value",0)), "frac", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:20], d.get("parity_check",{}).get("ok"))
except Exception as e:
    print("${wl} failed", e)
PY
done
tail -5 gpurun_out/${T}_bench.err
python __graft_entry__.py --smoke 2>&1 | tail -4
