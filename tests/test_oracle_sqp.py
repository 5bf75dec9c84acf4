"""The C restatement of the SQP outer loop (oracle/oracle_sqp.c, with the NLP generated as C from the .nl DAG) against the numpy
mirror of the product (restartsqp_b200/sqp_driver.py on the oracle-backed twin backend), instance by instance in single-instance
mode (the batched driver shares its init/hotstart state machine across the batch, the reference and the C oracle do not):
bitwise on the polynomial problems, to tolerance where libm and numpy evaluate pow/exp differently."""
import os

import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from oracle import oracle_py as orc
from oracle_backend import OracleQPInterface
from test_hs_suite import HS_DIR, perturbed_starts, F_STAR

POLYNOMIAL = ["hs015", "hs043", "hs071", "hs087", "hs099", "hs106", "hs113", "hs118"]


@pytest.mark.parametrize("name", POLYNOMIAL)
def test_c_oracle_equals_numpy_mirror_single_instance(name):
    nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    X = perturbed_starts(nlp, 4, 1)
    res_c = orc.SqpOracle(nlp, r.Options(iter_max=150)).solve_batch(X)
    for b in range(X.shape[0]):
        o1 = r.Options(iter_max=150)
        mk = lambda info, qt: r.QPhandler(info, qt, o1, batch=1, backend=OracleQPInterface(info, qt, o1, batch=1), refresh_ubA=True)
        try:
            r1 = BatchedSQP(nlp, x0=X[b:b + 1], options=o1, make_handler=mk).Optimize()
        except (r.QP_NOT_OPTIMAL, r.LP_NOT_OPTIMAL):
            continue  # batch == 1 keeps the reference's exceptions; the C oracle records the status instead
        assert int(r1.exitflag[0]) == int(res_c["exitflag"][b])
        assert int(r1.iters[0]) == int(res_c["iters"][b]) and int(r1.qp_iter[0]) == int(res_c["qp_iter"][b])
        if np.isfinite(r1.x[0]).all():
            assert np.array_equal(r1.x[0], res_c["x"][b]) and r1.obj[0] == res_c["obj"][b]


@pytest.mark.parametrize("name", sorted(F_STAR))
def test_c_oracle_reaches_known_optimum(name):
    nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    res = orc.SqpOracle(nlp, r.Options()).solve_batch(perturbed_starts(nlp, 3, 0), nthreads=2)
    assert (res["exitflag"] == 0).all()
    assert abs(res["obj"][0] - F_STAR[name]) <= 1e-3 * max(1.0, abs(F_STAR[name]))


@pytest.mark.parametrize("name", ["hs006", "hs015", "hs043", "hs113"])
def test_c_oracle_second_order_correction_equals_numpy_mirror(name):
    """second_order_correction (src/Algorithm.cpp:1140-1211, opt-in) in oracle_sqp.c against the numpy mirror, single instance,
    every QP / LP of both by the C oracle of the backend: identical exit flags, outer and QP iteration counts and iterates; the
    correction must actually be taken somewhere (the QP iteration counts differ from a run without it)."""
    nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    X = perturbed_starts(nlp, 6, 2)
    res_c = orc.SqpOracle(nlp, r.Options(iter_max=150, second_order_correction=True)).solve_batch(X)
    res_off = orc.SqpOracle(nlp, r.Options(iter_max=150)).solve_batch(X)
    compared = 0
    for b in range(X.shape[0]):
        o1 = r.Options(iter_max=150, second_order_correction=True)
        mk = lambda info, qt: r.QPhandler(info, qt, o1, batch=1, backend=OracleQPInterface(info, qt, o1, batch=1), refresh_ubA=True)
        try:
            r1 = BatchedSQP(nlp, x0=X[b:b + 1], options=o1, make_handler=mk).Optimize()
        except (r.QP_NOT_OPTIMAL, r.LP_NOT_OPTIMAL):
            continue
        assert int(r1.exitflag[0]) == int(res_c["exitflag"][b])
        assert int(r1.iters[0]) == int(res_c["iters"][b]) and int(r1.qp_iter[0]) == int(res_c["qp_iter"][b])
        if np.isfinite(r1.x[0]).all():
            assert np.array_equal(r1.x[0], res_c["x"][b]) and r1.obj[0] == res_c["obj"][b]
        compared += 1
    assert compared >= 3
    if name in ("hs006", "hs043"):
        assert (res_c["qp_iter"] != res_off["qp_iter"]).any()


def test_qp_unchanged_guard_of_setupQP():
    """Algorithm::setupQP throws QP_UNCHANGED when no Update_* flag was raised since the last solve (src/Algorithm.cpp:651-670): a
    rejected step whose ratio neither shrinks nor grows the radius (possible when the predicted reduction is not positive).  hs105
    gets there after its first iteration from most perturbed starts.  The C oracle and the numpy mirror end such an instance with
    Exitflag.QP_UNCHANGED (7, not a value of the reference's enum) instead of re-solving the same QP until iter_max."""
    nlp = AmplNLP(os.path.join(HS_DIR, "hs105.nl"))
    X = perturbed_starts(nlp, 8, 1)
    res_c = orc.SqpOracle(nlp, r.Options(iter_max=50)).solve_batch(X)
    unch = res_c["exitflag"] == int(r.Exitflag.QP_UNCHANGED)
    assert unch.sum() >= 4 and (res_c["iters"][unch] >= 1).all() and (res_c["iters"][unch] < 50).all()
    agree = 0
    for b in np.nonzero(unch)[0][:4]:
        o1 = r.Options(iter_max=50)
        mk = lambda info, qt: r.QPhandler(info, qt, o1, batch=1, backend=OracleQPInterface(info, qt, o1, batch=1), refresh_ubA=True)
        r1 = BatchedSQP(nlp, x0=X[b:b + 1], options=o1, make_handler=mk).Optimize()
        agree += int(r1.exitflag[0]) == int(r.Exitflag.QP_UNCHANGED) and int(r1.iters[0]) == int(res_c["iters"][b])
    assert agree >= 3  # libm and numpy evaluate hs105's exp / log differently in the last bits: not every start has to agree


@pytest.mark.parametrize("name", ["hs099", "hs108"])
def test_qp_failure_inside_the_penalty_loop_runs_the_rest_of_the_iteration(name):
    """QP_NOT_OPTIMAL inside update_penalty_parameter only leaves its while loop (src/Algorithm.cpp:932-935, 958-961): the
    acceptance test then sees the objective of an unsolved QP (getObjVal() = INFTY) and takes its failure branch, and Optimize
    still runs the trial point, the ratio test, iter++ and check_optimality before its loop condition ends the solve.  Instances
    that take this path (found with the oracle's counter): C oracle == numpy mirror on the oracle-backed backend, bit for bit."""
    import ctypes as C
    L = orc.lib()
    L.orc_sqp_penalty_qp_failures.restype = C.c_longlong
    nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    X = perturbed_starts(nlp, 64, 4)
    so = orc.SqpOracle(nlp, r.Options(iter_max=150))
    hit = compared = 0
    for b in range(X.shape[0]):
        L.orc_sqp_penalty_qp_failures()
        res_c = so.solve_batch(X[b:b + 1], nthreads=1)
        if L.orc_sqp_penalty_qp_failures() == 0:
            continue
        hit += 1
        assert 20 < int(res_c["exitflag"][0]) <= 30 or int(res_c["exitflag"][0]) == 0  # the QP's error code, or OPTIMAL by check_optimality
        o1 = r.Options(iter_max=150)
        mk = lambda info, qt: r.QPhandler(info, qt, o1, batch=1, backend=OracleQPInterface(info, qt, o1, batch=1), refresh_ubA=True)
        try:
            r1 = BatchedSQP(nlp, x0=X[b:b + 1], options=o1, make_handler=mk).Optimize()
        except (r.QP_NOT_OPTIMAL, r.LP_NOT_OPTIMAL):
            continue
        assert int(r1.exitflag[0]) == int(res_c["exitflag"][0]) and int(r1.iters[0]) == int(res_c["iters"][0])
        assert int(r1.qp_iter[0]) == int(res_c["qp_iter"][0])
        if np.isfinite(r1.x[0]).all():
            assert np.array_equal(r1.x[0], res_c["x"][0])
        compared += 1
    assert hit >= 1 and compared >= 1
