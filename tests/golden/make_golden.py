"""Regenerates the fixtures under tests/golden/ (golden ANSWERS) and altro_b200/data/ (problem INPUT
data) from the reference checkout.

Run HERE (the container that has /root/reference); the GPU box only sees the committed outputs.
  altro_b200/data/scotty_ref.json <- test/scotty.json (reference trajectory read by test_utils.cpp:240-289)
  scotty_mpc.json   <- test/scotty_mpc.json  (golden output of bicycle_test.cpp:247-359:
                                              200 warm-started MPC solves)
Only data is transferred (numbers), re-serialised compactly; no reference source is copied.
"""
import json
import os

REF = "/root/reference/test"
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "..", "altro_b200", "data")


def main():
    ref = json.load(open(os.path.join(REF, "scotty.json")))
    out = {"N": ref["N"], "tf": ref["tf"], "state_trajectory": ref["state_trajectory"],
           "input_trajectory": ref["input_trajectory"]}
    json.dump(out, open(os.path.join(DATA, "scotty_ref.json"), "w"), separators=(",", ":"))
    mpc = json.load(open(os.path.join(REF, "scotty_mpc.json")))
    out = {k: mpc[k] for k in ("N", "tf", "state_trajectory", "input_trajectory", "solve_iters",
                               "tracking_error")}
    json.dump(out, open(os.path.join(HERE, "scotty_mpc.json"), "w"), separators=(",", ":"))
    print("wrote scotty_ref.json, scotty_mpc.json")


if __name__ == "__main__":
    main()
