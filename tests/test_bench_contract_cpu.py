"""bench.py contract on the CPU: the reference arm (--impl reference) runs without a GPU and must
print ONE JSON line with the keys the driver reads; the per-kernel byte model must match the
DESIGN.md 'Kernels' table for the bench shape."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--batch", "64", "--ref-step-seconds", "0.3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "trajectory_solves_per_sec" and d["unit"] == "solves/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_kernel_byte_model_matches_design_table():
    sys.path.insert(0, ROOT)
    import bench
    from altro_b200 import problems as PR
    P = PR.bicycle(B=32, N=100, n=5)
    it = np.full(P.B, 10)
    ev = np.full(P.B, 30)
    m = bench.kernel_models(P, it, ev)
    # DESIGN.md section 4, (5,2) with the packed Jacobian (V = 15), bytes per trajectory-knot
    assert 8 * m["backward"]["doubles"] == 784        # 904 in a problem's first iteration
    assert m["backward"]["extra_bytes"] == (904 - 784) * P.B * 100
    assert 8 * m["fwd_rollout"]["doubles"] == 272
    # expansion (288 B per knot, N + 1 knots) and d(phi) scan (272 B per knot) of every iteration: done
    # inside the rollout passes by the follower warp, booked there
    assert m["fwd_rollout"]["extra_bytes"] == (288 * 101 + 272 * 100) * P.B * 10
    assert 8 * m["expand"]["doubles"] == 288
    assert 8 * m["fwd_criteria"]["doubles"] == 360 + 368
    assert m["backward"]["units"] == P.B * 10 * 100 and m["fwd_rollout"]["units"] == P.B * 20 * 100
    # k_phase_forward carries the sum of its sub-phases
    assert m["forward"]["extra_bytes"] == sum(8.0 * m[k]["doubles"] * m[k]["units"] + m[k].get("extra_bytes", 0.0)
                                              for k in ("fwd_rollout", "fwd_criteria"))
    # D(5,2) = 1552 bytes per knot-point-iteration
    assert 8 * (4 * 25 + 4 * 10 + 8 * 5 + 7 * 2) == 1552
