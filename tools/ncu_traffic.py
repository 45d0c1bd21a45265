#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full summaries (tools/ncu_summary.py output): DRAM bytes per
trajectory-knot of each phase kernel = (dram read + write of the launch) / (groups in the launch
* 32 problems * knots).  Usage: python tools/ncu_traffic.py workload N summary.json [...]"""
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
PHASE = {"k_phase_backward": "backward", "k_phase_forward": "forward", "k_phase_expand": "expand"}


def val(s):
    v, u = s.split()
    return float(v) * UNIT[u]


def main():
    wl, N = sys.argv[1], int(sys.argv[2])
    out = {}
    for path in sys.argv[3:]:
        for rep, launches in json.load(open(path)).items():
            for e in launches:
                name = next((p for p in PHASE if p in e["Kernel Name"]), None)
                if not name:
                    continue
                grid = [int(x) for x in re.findall(r"\d+", e["Grid Size"])]
                block = [int(x) for x in re.findall(r"\d+", e["Block Size"])]
                if block[0] == 128 and len(grid) > 1 and grid[1] > 1:   # knot-parallel kernels
                    units = grid[0] * 128 * grid[1]
                else:                                                    # one CTA per group
                    units = grid[0] * 32 * N
                b = val(e["dram__bytes_read.sum"]) + val(e["dram__bytes_write.sum"])
                d = out.setdefault(PHASE[name], {"dram_bytes": 0.0, "units": 0.0, "launches": 0, "kernels": []})
                d["dram_bytes"] += b
                d["units"] += units
                d["launches"] += 1
                if name not in d["kernels"]:
                    d["kernels"].append(name)
    for d in out.values():
        d["dram_bytes_per_unit"] = d["dram_bytes"] / d["units"]
        d["dram_bytes_per_launch"] = d["dram_bytes"] / d["launches"]  # captured launches: full batch
    print(json.dumps({wl: out}, indent=1))


if __name__ == "__main__":
    main()
