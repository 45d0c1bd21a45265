"""JSON trajectory I/O of the C ABI (include/altro_b200.h section D) against the reference's files:
the reader accepts test/scotty.json / test/scotty_mpc.json as they are (keys of
test/test_utils.cpp:240-289 and test/bicycle_test.cpp:344-359); the writer's single-problem layout
reads back identically with a stock JSON parser; batch files round-trip bit for bit."""
import json
import os

import numpy as np
import pytest

from altro_b200 import trajectory_io as TIO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reads_reference_trajectory_file():
    path = os.path.join(ROOT, "altro_b200", "data", "scotty_ref.json")
    d = TIO.read_trajectory(path)
    ref = json.load(open(path))
    assert d["batch"] == 1 and d["N"] == ref["N"] == 501 and d["tf"] == ref["tf"]
    assert np.array_equal(d["state_trajectory"][0], np.array(ref["state_trajectory"]))
    assert np.array_equal(d["input_trajectory"][0], np.array(ref["input_trajectory"]))
    assert d["solve_iters"] is None and d["tracking_error"] is None


def test_reads_reference_mpc_output_file():
    path = os.path.join(ROOT, "tests", "golden", "scotty_mpc.json")
    d = TIO.read_trajectory(path)
    ref = json.load(open(path))
    assert d["N"] == 200 and d["state_trajectory"].shape == (1, 201, 4) and d["input_trajectory"].shape == (1, 200, 2)
    assert np.array_equal(d["solve_iters"][0], np.array(ref["solve_iters"]))
    assert np.array_equal(d["tracking_error"][0], np.array(ref["tracking_error"]))
    assert d["solve_iters"].sum() == 627


def test_single_problem_writer_uses_the_reference_keys(tmp_path):
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "scotty_mpc.json")))
    out = tmp_path / "mpc.json"
    TIO.write_trajectory(out, ref["N"], ref["tf"], np.array(ref["state_trajectory"]),
                         np.array(ref["input_trajectory"]), ref["solve_iters"], ref["tracking_error"])
    back = json.load(open(out))                  # a stock parser reads what the reference would read
    assert set(back) == {"N", "tf", "state_trajectory", "input_trajectory", "solve_iters", "tracking_error"}
    for k in back:
        assert back[k] == ref[k], k


def test_batch_round_trip_is_exact(tmp_path):
    rng = np.random.default_rng(0)
    B, N, n, m = 5, 7, 4, 2
    X = rng.normal(size=(B, N + 1, n)) * 10.0 ** rng.integers(-8, 8, size=(B, N + 1, n))
    U = rng.normal(size=(B, N, m))
    it = rng.integers(1, 30, size=(B, N)).astype(np.int32)
    te = np.abs(rng.normal(size=(B, N)))
    out = tmp_path / "batch.json"
    TIO.write_trajectory(out, N, 0.7, X, U, it, te)
    assert json.load(open(out))["batch"] == B
    d = TIO.read_trajectory(out)
    assert d["batch"] == B and d["N"] == N and d["tf"] == pytest.approx(0.7, rel=1e-7)
    assert np.array_equal(d["state_trajectory"], X) and np.array_equal(d["input_trajectory"], U)
    assert np.array_equal(d["solve_iters"], it) and np.array_equal(d["tracking_error"], te)


def test_malformed_files_are_reported(tmp_path):
    from altro_b200.solver import AltroB200Error, ErrorCodes
    with pytest.raises(AltroB200Error) as e:
        TIO.read_trajectory(tmp_path / "missing.json")
    assert e.value.code == ErrorCodes.FileError
    bad = tmp_path / "bad.json"
    bad.write_text('{"N": 3, "state_trajectory": [[1, 2], [3]]}')      # ragged
    with pytest.raises(AltroB200Error) as e:
        TIO.read_trajectory(bad)
    assert e.value.code == ErrorCodes.DimensionMismatch
    bad.write_text('{"N": 3, "state_trajectory": [[1, 2], [3, 4]')     # truncated
    with pytest.raises(AltroB200Error) as e:
        TIO.read_trajectory(bad)
    assert e.value.code == ErrorCodes.FileError
