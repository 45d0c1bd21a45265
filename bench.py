#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched AL-iLQR solve path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, 1 rank/GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W     (CPU arm: oracle, all cores)

Metric (BASELINE.json): trajectory solves/sec, full AL-iLQR solve to the reference's convergence
criteria.  Workload at every N: BASELINE config "bicycle (n=5, m=2, N=100), batch 16384 random
goals" PER GPU (weak scaling; the batch shards embarrassingly, no collective on the data path).
A step = one batched Solve() of the whole per-GPU batch.

Prints ONE JSON line (rank 0).  `value` = solves/s with inputs resident in HBM; `e2e` = the same
through the public C ABI with host buffers (H2D of problem data and D2H of the solution inside
the timed region); `roofline` = the dominant phase kernel of the step: its algorithmic HBM bytes
per launch / its CUDA-event time, measured live in one extra solve with
every launch bracketed by events, against the measured copy bandwidth (MEASURED_PEAKS.json, else
the 6.65 TB/s fallback), `traffic` = its ncu DRAM bytes per launch (profiles/traffic.json);
`kernels` = the same for every phase; `cpu_baseline` = the CPU oracle (restatement of the
reference, oracle/) timed on this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from altro_b200 import problems as PR  # noqa: E402

METRIC = "trajectory_solves_per_sec"
UNIT = "solves/s"


def workload(name, per_gpu_batch, rank, world):
    """The per-rank slice of the global synthetic batch (rank r owns problems [r*B, (r+1)*B))."""
    B = per_gpu_batch
    gen = {
        "bicycle": lambda total: PR.bicycle(B=total, N=100, n=5),
        "pendulum": lambda total: PR.pendulum(B=total, N=100),
        "scotty": lambda total: PR.scotty(B=total, N=50, n=5),
        "scotty_mpc": lambda total: PR.scotty(B=total, N=50, n=5, margin=64),
        "chain12": lambda total: PR.chain(B=total, n=12, m=4, N=200),
        "chain6": lambda total: PR.chain(B=total, n=6, m=2, N=200),
    }.get(name)
    if gen is None and name.startswith("chain-"):   # chain-n-m-N: any shape of the BASELINE sweep
        cn, cm, cN = [int(v) for v in name.split("-")[1:4]]
        gen = lambda total: PR.chain(B=total, n=cn, m=cm, N=cN)
    P = gen(B * world)
    return P.subset(rank * B, (rank + 1) * B)


def workload_desc(name, P, world):
    return {"workload": f"{name}: {P.name}, n={P.n} m={P.m} N={P.N}, batch {P.B}/GPU x {world} GPU, "
                        f"synthetic random problems (SURVEY 8d seeds), iterations_max="
                        f"{P.options.get('iterations_max', 200)}, "
                        f"{'backtracking' if P.options.get('use_backtracking_linesearch') else 'cubic'} "
                        f"line search",
            "batch_per_gpu": P.B, "n": P.n, "m": P.m, "horizon": P.N,
            "timing": "CUDA events on the launching stream; max over ranks",
            "l2_policy": "working set (>= 1.4 GB of per-trajectory state) is far larger than the "
                         "126 MB L2; no flush needed"}


def algorithmic_bytes(P, iters, evals):
    """SURVEY.md 8(d) 'Algorithmic bytes per unit of work', summed over the batch.
    Unit = knot-point-iteration: 8*(4n^2+4nm+8n+7m) bytes; every merit evaluation beyond one per
    iteration adds (mn+3n+4m+(n+m)) doubles per knot; problem I/O (N+1)n+Nm doubles in and out.
    Constraint duals add 2 doubles per row per knot-iteration."""
    n, m, N = P.n, P.m, P.N
    D = 4 * n * n + 4 * n * m + 8 * n + 7 * m
    E = m * n + 3 * n + 4 * m + (n + m)
    rows = P.n_constraint_rows()
    it = iters.astype(np.float64)
    extra = np.maximum(evals.astype(np.float64) - it, 0.0)
    per = 8.0 * (N * it * (D + 2 * rows) + N * extra * E + 2 * ((N + 1) * n + N * m))
    return float(per.sum())


# Per-kernel split of the algorithmic bytes (DESIGN.md "Kernels"): compulsory HBM doubles per
# trajectory-knot each phase kernel must move, and how many trajectory-knots the ALGORITHM needs
# it to process in one solve (speculative / masked-off work is NOT counted as useful).
def kernel_models(P, iters, evals):
    n, m, N = P.n, P.m, P.N
    rows = P.n_constraint_rows()
    it = float(iters.astype(np.float64).sum())                       # backward passes
    roll = float(np.maximum(evals.astype(np.float64) - iters, 0).sum())  # rollouts the search consumed
    # Jacobian entries kept in HBM: dense n^2 + nm unless the model packs them (models.cuh, JacPack)
    jac = {PR.MODEL_BICYCLE5: 15, PR.MODEL_BICYCLE4: 12}.get(P.model_id, n * n + n * m)
    # sub-phases of k_phase_forward.  ALGORITHMIC bytes = the SURVEY 8(d) split of the reference
    # formulation, unchanged since round 1 so that the fractions stay comparable: per consumed
    # rollout r [xbar ubar q r c K d] (+ z) w x,u; per iteration one expansion (r x,u,q,r  w J,lx,lu
    # (+ z, z_est)) and one d(phi) scan (r K,d,J,lx,lu); costates + residuals + copy
    # (r x,xbar,P,p  w y;  r x,u,y,y+,J,lx,lu  w xbar,ubar).  The round-2 kernel moves less than that
    # (the follower warp needs no re-reads, goal-type costs do not stream
    # [q r c], the post-search pass reads every block once); the time of the expansion and of the
    # scan is inside the rollout passes now, so their bytes are booked there.
    roll_d = 2 * (n + m) + 1 + m * n + m + (n + m) + rows
    expand_d = 2 * (n + m) + (jac + n + m) + 2 * rows
    dphi_d = m * n + m + jac + n + m
    sub = {
        "fwd_rollout": dict(doubles=roll_d, units=roll * N,
                            extra_bytes=8.0 * (expand_d * it * (N + 1) + dphi_d * it * N)),
        "fwd_criteria": dict(doubles=(2 * n + n * n + n) + n + (2 * n + jac + 2 * (n + m)) + (n + m) + rows,
                             units=it * (N + 1)),
    }
    fwd_bytes = sum(8.0 * v["doubles"] * v["units"] + v.get("extra_bytes", 0.0) for v in sub.values())
    models = {
        # sweep: r J,lx,lu  w K,d,P,p (+ z_est rows).  scan: r q,r,c,K,d,x,u,J  w lx,lu (+ duals);
        # unconstrained problems after their first iteration scan only K,d,J,lx,lu and write nothing
        "backward": dict(kernel="k_phase_backward (Riccati sweep + alpha=0 scan)",
                         doubles=(jac + n + m) + (m * n + m + n * n + n) + rows
                         + ((2 * (n + m) + 1 + m * n + m + jac) + (n + m) + 2 * rows if rows
                            else (m * n + m + jac + n + m)),
                         units=it * N,
                         extra_bytes=0.0 if rows else 8.0 * (n + m + 1 + n + m) * P.B * N),
        "forward": dict(kernel="k_phase_forward (line search: rollout + follower + speculating warps, "
                               "state machines; fused expansion / costate / residual / copy pass, AL update)",
                        doubles=0.0, units=0.0, extra_bytes=fwd_bytes),
        "expand": dict(kernel="k_phase_expand (prologue: Jacobians, projected duals, gradients)",
                       doubles=2 * (n + m) + (jac + n + m) + 2 * rows, units=P.B * (N + 1)),
    }
    for k, v in sub.items():
        models[k] = dict(kernel=f"k_phase_forward / {k[4:]}", **v)
    return models


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active," \
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                                  ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload_name, phase):
    """dram read+write bytes per launch of a phase kernel from the committed ncu --set full capture
    of full-batch launches of this workload (profiles/traffic.json, tools/ncu_traffic.py)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload_name, {}).get(phase, {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def step_traffic(workload_name, kernels, ms_per_step, peak):
    """DRAM bytes the step actually moves: ncu dram read+write per full-batch launch of each kernel
    (profiles/traffic.json, captured with the sub-batch split off) x its launches per step, against
    the step time -- how close the step is to the memory system's limit on the traffic it generates
    (as opposed to `step_roofline`, which counts only the bytes the algorithm needs)."""
    tot, used = 0.0, []
    for ph in ("backward", "forward", "expand"):
        t = ncu_traffic(workload_name, ph)
        if t and ph in kernels:
            tot += t * kernels[ph]["launches"]
            used.append(ph)
    if not used:
        return None
    gbs = tot / (ms_per_step * 1e-3) / 1e9
    return {"dram_bytes_per_step": tot, "achieved": gbs, "unit": "GB/s", "frac_of_peak": gbs / peak,
            "kernels": used, "source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what
    the CPU arm is meant to measure, so the count is taken from the affinity mask)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(P, target_seconds, threads=None):
    """Times the CPU oracle on a bounded prefix of the workload.  Returns dict for the JSON."""
    from oracle import oracle as O
    threads = threads or host_threads()
    probe = min(P.B, 4 * threads)
    r = O.solve_batch(P, 0, probe, nthreads=threads)
    rate = probe / max(r["seconds"], 1e-9)
    nb = int(min(P.B, max(probe, rate * target_seconds)))
    r = O.solve_batch(P, 0, nb, nthreads=threads)
    return {"value": nb / r["seconds"], "unit": UNIT, "cores": int(threads), "kind": "port",
            "sample": f"first {nb} problems of the per-GPU batch, Solve() only, {r['seconds']:.2f} s "
                      f"wall, mean {r['iters'].mean():.1f} iterations / {r['merit_evals'].mean():.1f} "
                      f"merit evaluations per solve, oracle = C restatement of the reference "
                      f"(Eigen reference not buildable here: DESIGN.md)"}, r, nb


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (oracle port)."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    P = workload(args.workload, args.batch, 0, 1)
    threads = host_threads()
    probe = min(P.B, 4 * threads)
    r = O.solve_batch(P, 0, probe, nthreads=threads)
    rate = probe / max(r["seconds"], 1e-9)
    per_step = int(min(P.B, max(probe, rate * args.ref_step_seconds)))
    for _ in range(args.warmup):
        O.solve_batch(P, 0, min(per_step, 2 * threads), nthreads=threads)
    tot_s, tot_n, its = 0.0, 0, []
    for s in range(args.steps):
        b0 = (s * per_step) % max(P.B - per_step, 1)
        r = O.solve_batch(P, b0, b0 + per_step, nthreads=threads)
        tot_s += r["seconds"]
        tot_n += per_step
        its.append(r["iters"].mean())
    val = tot_n / tot_s
    sample = f"{per_step} problems of the per-GPU batch per step, {args.steps} steps, Solve() only, " \
             f"mean {np.mean(its):.1f} iterations per solve"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": workload_desc(args.workload, P, 1),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(threads), "kind": "port",
                            "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bicycle")
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--split", type=int, default=0, help="pipelined sub-batches (0 = automatic)")
    ap.add_argument("--slots", type=int, default=None, help="candidate steps per line-search round")
    ap.add_argument("--store", type=int, default=None, help="speculative candidates that keep their trajectory")
    ap.add_argument("--mpc-steps", type=int, default=10, choices=range(1, 65),
                    help="scotty_mpc: receding-horizon solves per bench step (all on the device)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {"bicycle": 16384, "pendulum": 4096, "scotty": 8192, "scotty_mpc": 8192, "chain12": 4096,
                      "chain6": 32768}.get(args.workload, 32768)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import altro_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solve path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    P = workload(args.workload, args.batch, rank, world)
    solver = altro_b200.make_solver(P, device=local_rank, nslots=args.slots)
    solver.SetStream(torch.cuda.current_stream().cuda_stream)
    solver.SetPipelineSplit(args.split)
    if args.store is not None:
        solver.SetCandidateStore(args.store)
    B, N, n, m = P.B, P.N, P.n, P.m
    # scotty_mpc: a bench step is `mpc_steps` consecutive receding-horizon solves of every problem,
    # the MPC step between two solves (x0 <- x_[1], ShiftTrajectory, window + 1 with the reference's
    # cost update) done by altro_b200_mpc_step on the device: no host round trip inside the step
    mpc_steps = args.mpc_steps if args.workload == "scotty_mpc" else 1
    if mpc_steps > 1:
        u0 = P.U0[0, 0]
        solver.SetMpcCostUpdate(1, float(0.5 * u0 @ (P.Rd[0] * u0)))
    solves_per_step = mpc_steps

    # ------------------------------------------------------------ (1) resident-in-HBM steps
    def resident_step(ev=None):
        solver.ResetTrajectory()
        solver.ResetDuals()
        if mpc_steps > 1:
            solver.SetInitialState(P.x0)
        if ev:
            ev[0].record()
        solver.SolveAsync()
        for _ in range(mpc_steps - 1):
            solver.MpcStep()
            solver.SolveAsync()
        if ev:
            ev[1].record()
        if mpc_steps > 1:                # window back to where it started (q, r, c of offset 0)
            solver.AdvanceWindow(-(mpc_steps - 1))

    for _ in range(args.warmup):
        resident_step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    kernel_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                 for _ in range(args.steps)]
    l0 = solver.KernelLaunches()
    barrier()
    clocks.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for s in range(args.steps):
        resident_step(kernel_ev[s])
    t_end.record()
    barrier()
    clk = clocks.stop()
    launches = solver.KernelLaunches() - l0
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_ev]))
    iters, evals, status = solver.GetIterations(), solver.GetMeritEvals(), solver.GetStatus()

    # ------------------------------------------------------------ (2) end to end through the C ABI
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    x0_h, xref_h, uref_h = pin(P.x0), pin(P.xref) if P.ref_mode == PR.REF_GOAL else None, \
        pin(P.uref) if P.ref_mode == PR.REF_GOAL else None
    U0_h = pin(P.U0)
    X_h = torch.empty((B, N + 1, n), dtype=torch.float64).pin_memory().numpy()
    U_h = torch.empty((B, N, m), dtype=torch.float64).pin_memory().numpy()
    st_h = torch.empty((B,), dtype=torch.int32).pin_memory().numpy()
    phi_h = torch.empty((B,), dtype=torch.float64).pin_memory().numpy()

    def e2e_step():
        solver.SetInitialState(x0_h)
        if P.ref_mode == PR.REF_GOAL:
            solver.SetLQRCost(P.Qd[0], P.Rd[0], xref_h, uref_h, 0, N)
            solver.SetLQRCost(P.Qd[N], P.Rd[N - 1], xref_h, uref_h, N, N + 1)
        else:
            altro_b200.solver.set_cost(solver, P)
        solver.SetInput(U0_h)
        solver.ResetDuals()
        solver.Solve()
        for _ in range(mpc_steps - 1):
            solver.GetInputs(out=U_h)        # the control an MPC loop applies comes back every step
            solver.MpcStep()
            solver.Solve()
        if mpc_steps > 1:
            solver.AdvanceWindow(-(mpc_steps - 1))
        solver.GetStates(out=X_h)
        solver.GetInputs(out=U_h)
        solver.GetStatus(out=st_h)
        solver.GetFinalObjective(out=phi_h)

    h2d = x0_h.nbytes + U0_h.nbytes
    if P.ref_mode == PR.REF_GOAL:
        h2d += 2 * (xref_h.nbytes + uref_h.nbytes)
    elif P.ref_mode == PR.REF_WINDOW:
        h2d += P.xref.nbytes + P.uref.nbytes + P.offsets.nbytes
    d2h = X_h.nbytes + U_h.nbytes * mpc_steps + st_h.nbytes + phi_h.nbytes
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for s in range(args.steps):
        e2e_step()
    e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)

    # ------------------------------------------------------------ (3) per-kernel times (rank 0)
    # one extra solve, sub-batch pipelining off and every launch bracketed by CUDA events on the
    # launching stream (the pipeline's own instrumentation): gives each phase kernel's launches
    # and summed duration without overlap from other sub-batches
    phase_stats = None
    if rank == 0:
        solver.SetPipelineSplit(1)
        solver.ResetTrajectory()
        solver.ResetDuals()
        solver.Solve()                       # warm (no split) instantiation
        solver.SetProfiling(1)
        solver.ResetTrajectory()
        solver.ResetDuals()
        solver.Solve()
        phase_stats, _ = solver.GetPhaseStats()
        solver.SetProfiling(0)
        solver.SetPipelineSplit(args.split)

    # ------------------------------------------------------------ max over ranks
    times = torch.tensor([elapsed_ms, e2e_ms, kernel_ms], dtype=torch.float64, device="cuda")
    counts = torch.tensor([float(B)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, kernel_ms_max = [float(v) for v in times.cpu()]
    total_solves = float(counts.cpu()[0])

    out = None
    if rank == 0:
        value = total_solves * solves_per_step * args.steps / (elapsed_ms * 1e-3)
        e2e_value = total_solves * solves_per_step * args.steps / (e2e_ms * 1e-3)
        peak, peak_src = peak_hbm()
        # scotty_mpc: the statistics are those of the step's LAST solve; the warm-started solves of
        # one step do similar work, so the step's bytes are taken as mpc_steps times that
        alg_bytes = algorithmic_bytes(P, iters, evals) * solves_per_step
        models = kernel_models(P, iters, evals)
        whole = ("init_rollout", "expand", "backward", "forward")   # launches; fwd_* are shares of forward
        tot_ms = sum(phase_stats[k]["ms"] for k in whole)
        kernels = {}
        for ph, st_ in phase_stats.items():
            if ph not in models or st_["launches"] == 0 or st_["ms"] <= 0:
                continue
            mdl = models[ph]
            bytes_total = 8.0 * mdl["doubles"] * mdl["units"] + mdl.get("extra_bytes", 0.0)
            kernels[ph] = {"kernel": mdl["kernel"], "launches": st_["launches"], "ms": st_["ms"],
                           "share_of_step": st_["ms"] / tot_ms,
                           "us_per_launch": 1e3 * st_["ms"] / st_["launches"],
                           "algorithmic_bytes_per_launch": bytes_total / st_["launches"],
                           "achieved_gbs": bytes_total / (st_["ms"] * 1e-3) / 1e9,
                           "frac": bytes_total / (st_["ms"] * 1e-3) / 1e9 / peak}
        dom = max((k for k in kernels if k in whole), key=lambda k: kernels[k]["ms"])
        kd = kernels[dom]
        traffic = ncu_traffic(args.workload, dom)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": workload_desc(args.workload, P, world),
               "clocks": clk,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "achieved": kd["achieved_gbs"], "peak": peak, "unit": "GB/s",
                            "frac": kd["frac"], "traffic": traffic,
                            "kernel": kd["kernel"], "launches_per_step": kd["launches"],
                            "kernel_us_per_launch": kd["us_per_launch"],
                            "share_of_step": kd["share_of_step"],
                            "algorithmic_bytes_per_launch": kd["algorithmic_bytes_per_launch"],
                            "peak_source": peak_src,
                            "note": "dominant kernel of the step; per-kernel algorithmic bytes = "
                                    "DESIGN.md 'Kernels' table x the trajectory-knots the algorithm "
                                    "needs (speculative candidates not counted); timed with CUDA "
                                    "events around every launch, sub-batch pipelining off; the "
                                    "fwd_* entries of `kernels` split k_phase_forward by its "
                                    "in-kernel %globaltimer sub-phase clocks"},
               "kernels": kernels,
               "step_roofline": {"algorithmic_bytes_per_step": alg_bytes,
                                 "achieved": alg_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9,
                                 "frac": alg_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9 / peak,
                                 "note": "SURVEY 8(d) bytes of the whole solve / step time"},
               "step_traffic": step_traffic(args.workload, kernels, elapsed_ms / args.steps, peak),
               "solve_stats": {"mean_iterations": float(iters.mean()),
                               "mean_merit_evals": float(evals.mean()),
                               "success_frac": float((status == 0).mean()),
                               "max_iterations_frac": float((status == 2).mean()),
                               "hbm_bytes_resident": int(solver.DeviceBytes())}}
        # solves that reach the reference's convergence criteria (status Success), per second
        out["converged_solves_per_sec"] = value * float((status == 0).mean())
        if not args.no_cpu_baseline and world == 1:
            cb, ref, nb = cpu_oracle_rate(P, args.cpu_seconds)
            out["cpu_baseline"] = cb
            # the same port on ONE host core ("on 1 core and on all cores")
            cb1, _, _ = cpu_oracle_rate(P, max(2.0, args.cpu_seconds / 3), threads=1)
            out["cpu_baseline_1core"] = {k: cb1[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if not args.no_check:
                from parity_util import compare, compare_all
                gpu = {"X": X_h, "U": U_h, "status": st_h, "iters": iters, "cost": phi_h}
                try:
                    rep = compare(gpu, ref)
                    out["parity_check"] = {"ok": True, **rep}
                except AssertionError as e:
                    out["parity_check"] = {"ok": False, "error": str(e)[:300]}
                # EVERY problem of the sample, converged or not, after 1 / 3 / 10 iterations: the
                # same iteration count and status, and how far states / inputs / cost are apart
                from oracle import oracle as O
                low = {}
                for itmax in (1, 3, 10):
                    Pl = P.subset(0, P.B)
                    Pl.options = dict(P.options, iterations_max=itmax)
                    solver.SetOptions(altro_b200.default_options(**Pl.options))
                    solver.ResetTrajectory()
                    solver.ResetDuals()
                    solver.Solve()
                    g = {"X": solver.GetStates(), "U": solver.GetInputs(), "status": solver.GetStatus(),
                         "iters": solver.GetIterations(), "cost": solver.GetFinalObjective()}
                    r = O.solve_batch(Pl, 0, min(nb, P.B), nthreads=host_threads())
                    low[str(itmax)] = compare_all(g, r)
                solver.SetOptions(altro_b200.default_options(**P.options))
                out["parity_check"]["all_problems_low_iteration"] = low
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    solver.close()


if __name__ == "__main__":
    main()
