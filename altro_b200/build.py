"""Builds altro_b200/libaltro_b200.so (sm_100a) in-tree with nvcc.

The solve kernel is instantiated per (model, n, m); each group is its own translation unit
(csrc/solve_inst.cu with -DALTRO_INST=g) compiled in parallel, then everything is linked into one
shared library exporting the C ABI of include/altro_b200.h.  Objects are cached under
altro_b200/build/ and rebuilt only when a source they depend on is newer.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libaltro_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: no implicit multiply-add contraction, so every kernel that evaluates the same
# expression (persistent kernel, phase kernels) produces the same bits; the matrix helpers use
# explicit fma().
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
# experiment knobs: CTA shape of the forward kernel (solver_phases.cuh)
for _k in ("ALTRO_FWD_THREADS", "ALTRO_FWD_CTAS"):
    if os.environ.get(_k):
        FLAGS += [f"-D{_k}={os.environ[_k]}"]
N_INST = 8
UNITS = [("capi", "capi.cu", []), ("tvlqr", "tvlqr.cu", []), ("tvlqr_batch", "tvlqr_batch.cu", []),
         ("facade", "altro_solver.cpp", []),
         ("json_io", "json_io.cpp", [])] + \
        [(f"solve_inst_{g}", "solve_inst.cu", [f"-DALTRO_INST={g}"]) for g in range(N_INST)]


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "..", "include", "altro_b200.h"),
         os.path.join(HERE, "..", "include", "altro", "altro_solver.hpp")]


def _compile(unit, verbose, skip=()):
    name, src, defs = unit
    srcp = os.path.join(CSRC, src)
    if not os.path.exists(srcp):
        return name, None, ""
    obj = os.path.join(OBJ, name + ".o")
    if name in skip:
        return name, obj, "stale (ALTRO_ONLY)"
    newest = max(os.path.getmtime(p) for p in [srcp] + _deps())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return name, obj, "cached"
    cmd = [NVCC] + FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stderr}")
    return name, obj, r.stderr


def build(verbose=False, jobs=None):
    """ALTRO_ONLY=2,1 (env) recompiles only those solve_inst groups (+capi/tvlqr) and links the rest
    from stale objects -- for quick kernel experiments; never commit a library built that way."""
    os.makedirs(OBJ, exist_ok=True)
    only = os.environ.get("ALTRO_ONLY")
    skip = set()
    if only:
        keep = {f"solve_inst_{g}" for g in only.split(",")} | {"capi", "tvlqr", "tvlqr_batch", "facade", "json_io"}
        # stale objects of the other groups are linked as they are (their mtime is left alone, so
        # the next full build recompiles them)
        skip = {name for name, _, _ in UNITS
                if name not in keep and os.path.exists(os.path.join(OBJ, name + ".o"))}
    jobs = jobs or min(len(UNITS), os.cpu_count() or 4)
    objs, logs = [], {}
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for name, obj, log in ex.map(lambda u: _compile(u, verbose, skip), UNITS):
            if obj:
                objs.append(obj)
                logs[name] = log
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if stale:
        cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return LIB, logs


if __name__ == "__main__":
    lib, logs = build(verbose="-v" in sys.argv)
    for k, v in logs.items():
        if v and v != "cached":
            print(f"==== {k}\n{v}")
    print("built", lib)
