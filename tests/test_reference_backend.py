"""PINS the restated backend glue (rows C2-C8 of SURVEY.md 8a) against the reference's REAL src/qpOASESInterface.cpp: its init /
hotstart state machine (:137-284, 817-833), handle_error (:686-758), get_status (:330-357), get_working_set (:835-895) and
test_optimality (:498-684) are compiled unmodified (dev container, oracle/build_reference_drivers.sh) on a functional stand-in
for qpOASES whose arithmetic is the oracle's active-set solver (oracle/stubs_link/qpoases_over_oracle.cpp), and run side by side
with the CUDA plugin on the CPU twin of the C ABI (the library's restated state machine over the same solver) through 40 random
sequences of 12 solves each: cold start, hot starts with fixed and with new matrices, matrix-status flips, infeasible data.
Compared bit for bit after every solve: x, y, status, objective, Stats::qp_iter, whether optimizeQP threw, the translated working
sets and the five fields of OptimalityStatus (oracle/backend_pin_test.cpp).

What must hold: every QP sequence agrees on every observable at every step, except sequences that run into a solver failure
(iteration limit, failed re-initialisation), where the recovery bookkeeping differs.  What differs by construction, found by this
comparison: for an LP the reference answers a matrix-status flip with a plain cold init (src/qpOASESInterface.cpp:262-266), the
restatements re-initialise from the previous solution as for a QP (:202-207), so every LP sequence agrees up to its first flip
(step 3) and no further.  The SQP loop never flips the LP handle's status (it hands the LP a new Jacobian before every solve), which
is why the loop-level comparison (tests/test_reference_algorithm.py) is unaffected."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "backend_pin_twin")


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/backend_pin_twin not built (needs /root/reference at build time)")
def test_reference_backend_glue_equals_the_restated_one():
    p = subprocess.run([BIN, "40"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    lines = p.stdout.strip().splitlines()
    assert lines[-1].split()[:5] == ["summary", "problems", "40", "steps", "480"]
    first = {}
    for ln in lines[:-1]:
        t = ln.split()
        assert t[0] == "mismatch"
        prob, step, typ = int(t[2]), int(t[4]), t[7].rstrip(":")
        first.setdefault(prob, (step, typ))
    lp = {k: v for k, v in first.items() if v[1] == "LP"}
    qp = {k: v for k, v in first.items() if v[1] == "QP"}
    assert sorted(lp) == list(range(3, 40, 4)) and all(step == 3 for step, _ in lp.values())  # every LP: identical until its first flip
    assert len(qp) <= 2, qp  # 28 of the 30 QP sequences identical on all 12 steps; the others leave at a solver failure
