#!/usr/bin/env python
"""profiles/r02_sass_summary.txt: for every kernel of the built objects (altro_b200/build/*.o) the
counts of the SASS mnemonics that prove what the kernel is made of -- UBLKCP (cp.async.bulk = TMA
bulk copies), SYNCS (mbarrier ops), DFMA/DMUL/DADD (FP64 pipe), MUFU, LDS/STS, LDG/STG, BAR -- plus
registers per thread, stack and shared memory from cuobjdump --dump-resource-usage.  No UTMALDG /
UTCMMA / HMMA is expected: the blocks are FP64 and at most 12 x 12 (BASELINE.json north_star).
Run HERE (cuobjdump needs no GPU):  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR",
        "UTMALDG", "UTCMMA", "HMMA"]


def short(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    out = out.replace("altro_b200::", "").replace("(anonymous namespace)::", "")
    return re.sub(r"\(.*", "", out)[:110]


def main():
    print(__doc__.split("Run HERE")[0])
    for obj in sorted(glob.glob(os.path.join(ROOT, "altro_b200", "build", "*.o"))):
        res = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                cur = m.group(1)
                continue
            m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
            if m and cur:
                usage[cur] = tuple(int(v) for v in m.groups())
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        parts = re.split(r"\n\s*Function : ", sass)
        if len(parts) < 2:
            continue
        print(f"== {os.path.basename(obj)}")
        for p in parts[1:]:
            name = p.split("\n")[0].strip()
            ops = collections.Counter()
            for line in p.split("\n"):
                m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
                if m:
                    ops[m.group(1)] += 1
            if not any(k in name for k in ("k_phase", "solve_kernel", "k_tvlqr", "k_knot", "k_calc", "k_open")):
                continue
            reg, stack, shared = usage.get(name, (0, 0, 0))
            cnt = " ".join(f"{k}={ops[k]}" for k in WANT if ops[k] or k in ("UBLKCP", "SYNCS", "DFMA", "UTMALDG", "UTCMMA"))
            print(f"  {short(name)}\n      regs={reg} stack={stack}B static_smem={shared}B insts={sum(ops.values())}  {cnt}")


if __name__ == "__main__":
    main()
