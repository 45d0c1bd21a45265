"""The N>1 path of bench.py on CPU: world_size-2 gloo processes, each owning its contiguous slice
of the synthetic batch (contiguous shards, no collective on the data path, only the
timing reduction and a host-side gather).  The compute of each shard runs on the CPU ORACLE here
(test infrastructure) because the product has no CPU path; what is tested is the sharding logic:
slices are disjoint, cover the batch, and the gathered result equals the unsharded one."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import torch, torch.distributed as dist
    import bench
    from oracle import oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B = 24
    P = bench.workload("pendulum", B, rank, world)           # this rank's shard
    r = O.solve_batch(P, nthreads=1)
    # timing reduction used by bench.py: max over ranks; count: sum over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([float(P.B)], dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    # host-side gather of the per-shard results
    xs = [torch.zeros((B, P.N + 1, P.n), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(xs, torch.from_numpy(r["X"]))
    x0s = [torch.zeros((B, P.n), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(x0s, torch.from_numpy(np.ascontiguousarray(P.x0)))
    if rank == 0:
        np.save(os.environ["OUT"], torch.cat(xs).numpy())
        np.save(os.environ["OUT"] + ".x0.npy", torch.cat(x0s).numpy())
        print(json.dumps({"tmax": float(t), "count": float(c)}))
    dist.barrier(); dist.destroy_process_group()
''')


def test_two_rank_gloo_sharding(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    out = str(tmp_path / "gathered.npy")
    env = dict(os.environ, OUT=out, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    info = json.loads(line)
    assert info == {"tmax": 2.0, "count": 48.0}
    import bench
    full = bench.workload("pendulum", 48, 0, 1)
    x0 = np.load(out + ".x0.npy")
    assert np.array_equal(x0, full.x0), "shards are not the contiguous slices of the global batch"
    ref = oracle.solve_batch(full, nthreads=2)
    assert np.array_equal(np.load(out), ref["X"])


def test_workload_slices_partition_the_batch():
    import bench
    for name in ("bicycle", "scotty"):
        whole = bench.workload(name, 64, 0, 1)
        parts = [bench.workload(name, 16, r, 4) for r in range(4)]
        assert np.array_equal(np.concatenate([p.x0 for p in parts]), whole.x0)
        if whole.offsets is not None:
            assert np.array_equal(np.concatenate([p.offsets for p in parts]), whole.offsets)
        else:
            assert np.array_equal(np.concatenate([p.xref for p in parts]), whole.xref)
