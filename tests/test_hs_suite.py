"""BASELINE.json configs[2]: the Hock-Schittkowski suite through the batched SQP outer loop, on CPU with the oracle-backed twin of
the QP backend (the host logic under test is the product's: nl_reader + sqp_driver + QPhandler).  Known optima from
Hock & Schittkowski, "Test Examples for Nonlinear Programming Codes" (1981)."""
import os

import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from oracle_backend import OracleQPInterface

HS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hs_nl")
F_STAR = {"hs001": 0.0, "hs012": -30.0, "hs024": -1.0, "hs035": 1.0 / 9.0, "hs043": -44.0, "hs048": 0.0, "hs051": 0.0,
          "hs071": 17.0140173, "hs076": -4.681818181, "hs100": 680.6300573, "hs113": 24.3062091, "hs117": 32.348679}


def perturbed_starts(nlp, B, k):
    """SURVEY.md 8d config 3"""
    x0, _ = nlp.Get_starting_point()
    xl, xu, _, _ = nlp.Get_bounds_info()
    rng = np.random.default_rng(71000 + k)
    X = np.clip(x0 * (1 + 0.1 * rng.standard_normal((B, nlp.n))) + 0.1 * rng.standard_normal((B, nlp.n)), xl, xu)
    X[0] = np.clip(x0, xl, xu)
    return X


@pytest.mark.parametrize("name", sorted(F_STAR))
def test_hs_problem_reaches_known_optimum(name):
    nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    B = 3
    opt = r.Options()
    mk = lambda info, qptype: r.QPhandler(info, qptype, opt, batch=B, backend=OracleQPInterface(info, qptype, opt, batch=B), refresh_ubA=True)
    res = BatchedSQP(nlp, x0=perturbed_starts(nlp, B, 0), options=opt, make_handler=mk).Optimize()
    assert (res.exitflag == int(r.Exitflag.OPTIMAL)).all(), res.exitflag
    # the nominal start reaches the tabulated optimum; perturbed starts of the non-convex problems may stop at another KKT point
    assert abs(res.obj[0] - F_STAR[name]) <= 1e-3 * max(1.0, abs(F_STAR[name]))
    assert np.isfinite(res.obj).all() and (res.KKT_error < 1e-3).all()


@pytest.mark.parametrize("name", ["hs006", "hs043", "hs100"])
def test_second_order_correction_on_rejected_steps(name):
    """src/Algorithm.cpp:1140-1211 (opt-in): these problems reject steps from the nominal start, so the corrected step is
    exercised; the run must still end OPTIMAL at the tabulated value and must not need more outer iterations than without it."""
    fstar = {"hs006": 0.0, "hs043": -44.0, "hs100": 680.6300573}[name]
    nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    B = 3
    X = perturbed_starts(nlp, B, 0)
    out = {}
    for soc in (False, True):
        opt = r.Options(second_order_correction=soc)
        mk = lambda info, qptype: r.QPhandler(info, qptype, opt, batch=B, backend=OracleQPInterface(info, qptype, opt, batch=B), refresh_ubA=True)
        out[soc] = BatchedSQP(nlp, x0=X, options=opt, make_handler=mk).Optimize()
        assert (out[soc].exitflag == int(r.Exitflag.OPTIMAL)).all()
        assert abs(out[soc].obj[0] - fstar) <= 1e-3 * max(1.0, abs(fstar))
    assert out[True].iters.sum() <= out[False].iters.sum()
