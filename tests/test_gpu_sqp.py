"""BASELINE.json config 1 / 3 on the GPU: HS071 through the batched SQP loop (mirror of src/Algorithm.cpp) with the
CUDA backend behind QPhandler, compared iterate-for-iterate with the same loop driven by the CPU oracle twin."""
import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.sqp_driver import BatchedSQP, HS071
from oracle_backend import OracleQPInterface

pytestmark = pytest.mark.gpu
X_STAR = np.array([1.0, 4.74299963, 3.82114998, 1.37940829])


def starts(B, seed=71000):
    rng = np.random.default_rng(seed)
    x0 = np.array([1.0, 5.0, 5.0, 1.0])
    s = np.clip(x0 * (1 + 0.1 * rng.standard_normal((B, 4))) + 0.1 * rng.standard_normal((B, 4)), 1.0, 5.0)
    s[0] = x0
    return s


def test_hs071_gpu_equals_oracle_driven_loop(gpu_lib):
    B = 16
    s = starts(B)
    opt_g, opt_o = r.Options(), r.Options()
    res_g = BatchedSQP(HS071(), x0=s, options=opt_g).Optimize()
    mk = lambda info, qt: r.QPhandler(info, qt, opt_o, batch=B, backend=OracleQPInterface(info, qt, opt_o, batch=B), refresh_ubA=True)
    res_o = BatchedSQP(HS071(), x0=s, options=opt_o, make_handler=mk).Optimize()
    assert (res_g.exitflag == res_o.exitflag).all() and (res_g.exitflag == int(r.Exitflag.OPTIMAL)).all()
    assert (res_g.iters == res_o.iters).all() and (res_g.qp_iter == res_o.qp_iter).all()
    assert np.abs(res_g.x - res_o.x).max() <= 1e-10  # the QP solutions are bit-identical, so are the iterates
    assert (res_g.rho == res_o.rho).all() and (res_g.delta == res_o.delta).all()
    assert np.abs(res_g.x - X_STAR).max() < 1e-4


def test_hs071_many_perturbed_starts(gpu_lib):
    B = 2000
    res = BatchedSQP(HS071(), x0=starts(B, seed=71001)).Optimize()
    ok = res.exitflag == int(r.Exitflag.OPTIMAL)
    assert ok.mean() > 0.99, np.unique(res.exitflag, return_counts=True)
    assert np.abs(res.x[ok] - X_STAR).max() < 1e-3
    assert np.abs(res.obj[ok] - 17.0140173).max() < 1e-3
