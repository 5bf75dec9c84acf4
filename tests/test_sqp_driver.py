"""The batched SQP outer loop (restartsqp_b200/sqp_driver.py, mirror of src/Algorithm.cpp) on CPU: the host logic is
driven through the oracle-backed twin of the QP backend (tests/oracle_backend.py).  BASELINE.json config 1: HS071."""
import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.sqp_driver import BatchedSQP, HS071, classify_single_constraint, BOUNDED, EQUAL, BOUNDED_ABOVE, BOUNDED_BELOW, UNBOUNDED
from oracle_backend import OracleQPInterface

X_STAR = np.array([1.0, 4.74299963, 3.82114998, 1.37940829])  # Hock-Schittkowski 71
F_STAR = 17.0140173


def oracle_handler_factory(batch, options):
    def make(info, qptype):
        return r.QPhandler(info, qptype, options, batch=batch, backend=OracleQPInterface(info, qptype, options, batch=batch),
                           refresh_ubA=True)
    return make


def test_classify_single_constraint_matches_reference():
    lo = np.array([[0.0, -1e19, 1.0, -1e19, 1.0]])
    hi = np.array([[1.0, 1.0, 1e19, 1e19, 1.0]])
    assert classify_single_constraint(lo, hi)[0].tolist() == [BOUNDED, BOUNDED_ABOVE, BOUNDED_BELOW, UNBOUNDED, EQUAL]


def test_hs071_single_instance_converges():
    opt = r.Options()
    alg = BatchedSQP(HS071(), options=opt, make_handler=oracle_handler_factory(1, opt))
    res = alg.Optimize()
    assert int(res.exitflag[0]) == int(r.Exitflag.OPTIMAL)
    assert np.abs(res.x[0] - X_STAR).max() < 1e-4
    assert abs(res.obj[0] - F_STAR) < 1e-4
    assert res.iters[0] < 50 and res.qp_iter[0] > 0


def test_hs071_batched_equals_single_runs():
    """Masked batched execution must reproduce, instance by instance, what independent single-instance runs give."""
    rng = np.random.default_rng(71000)
    x0 = np.array([1.0, 5.0, 5.0, 1.0])
    B = 6
    starts = np.clip(x0 * (1 + 0.1 * rng.standard_normal((B, 4))) + 0.1 * rng.standard_normal((B, 4)), 1.0, 5.0)
    starts[0] = x0
    opt = r.Options()
    resB = BatchedSQP(HS071(), x0=starts, options=opt, make_handler=oracle_handler_factory(B, opt)).Optimize()
    assert (resB.exitflag == int(r.Exitflag.OPTIMAL)).all()
    assert np.abs(resB.x - X_STAR).max() < 1e-4
    for b in range(B):
        opt1 = r.Options()
        res1 = BatchedSQP(HS071(), x0=starts[b:b + 1], options=opt1, make_handler=oracle_handler_factory(1, opt1)).Optimize()
        assert int(res1.exitflag[0]) == int(resB.exitflag[b])
        # identical unless another instance of the batch switched the shared hot-start mode (documented batch semantics)
        assert np.abs(res1.x[0] - resB.x[b]).max() < 1e-6


def test_second_order_correction_option():
    """src/Algorithm.cpp:1140-1211 (off by default): with the option on, rejected steps get a second QP around the trial point; the
    run still converges to the HS071 optimum, single-instance and batched runs agree, and with no rejection it changes nothing."""
    rng = np.random.default_rng(5)
    x0 = np.array([1.0, 5.0, 5.0, 1.0])
    B = 5
    starts = np.clip(x0 * (1 + 0.3 * rng.standard_normal((B, 4))) + 0.3 * rng.standard_normal((B, 4)), 1.0, 5.0)
    starts[0] = x0
    on, off = r.Options(second_order_correction=True), r.Options()
    res_on = BatchedSQP(HS071(), x0=starts, options=on, make_handler=oracle_handler_factory(B, on)).Optimize()
    res_off = BatchedSQP(HS071(), x0=starts, options=off, make_handler=oracle_handler_factory(B, off)).Optimize()
    assert (res_on.exitflag == int(r.Exitflag.OPTIMAL)).all() and (res_off.exitflag == int(r.Exitflag.OPTIMAL)).all()
    assert np.abs(res_on.x - X_STAR).max() < 1e-4
    assert res_on.qp_iter.sum() >= res_off.qp_iter.sum() - 50  # extra QPs only where a step was rejected
    for b in range(B):
        o1 = r.Options(second_order_correction=True)
        r1 = BatchedSQP(HS071(), x0=starts[b:b + 1], options=o1, make_handler=oracle_handler_factory(1, o1)).Optimize()
        assert int(r1.exitflag[0]) == int(res_on.exitflag[b]) and np.abs(r1.x[0] - res_on.x[b]).max() < 1e-6
