#!/usr/bin/env python
"""Source-level hot spots of a kernel from an ncu --set full --import-source on capture:
warp-stall samples and executed instructions per CUDA source line (top N), plus the instruction
mix by SASS opcode.  Usage: python tools/ncu_hotspots.py capture.ncu-rep [N] > profiles/...txt"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = page(rep, "--print-source", "sass,cuda")
    cur, agg, tot_s, tot_i, kernel = None, {}, 0, 0, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kernel = r[1]
        elif len(r) > 8 and r[2] == "-" and r[0].isdigit():
            try:
                s, i = int(r[4]), int(r[7])
            except ValueError:
                continue
            agg[(cur, int(r[0]))] = (s, i, r[1].strip())
            tot_s += s
            tot_i += i
    print(f"kernel: {kernel}\ncapture: {rep}\nwarp-stall samples {tot_s}, warp instructions executed {tot_i}\n")
    print(f"top {top} source lines by stall samples (share of samples | share of instructions | line)")
    for (f, l), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100 * s / tot_s:5.1f}%  {100 * i / tot_i:5.1f}%  {f}:{l}  {src[:90]}")
    sass = page(rep)
    hdr = next(r for r in sass if r and r[0] == "Address")
    ci = hdr.index("Instructions Executed")
    mix = collections.Counter()
    for r in sass:
        if len(r) <= ci or not r[0].startswith("0x"):
            continue
        ins = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip())
        try:
            mix[ins.split()[0].split(".")[0]] += int(r[ci])
        except (ValueError, IndexError):
            pass
    tot = sum(mix.values())
    print("\ninstruction mix (executed warp instructions by opcode):")
    print("  " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in mix.most_common(16)))


if __name__ == "__main__":
    main()
