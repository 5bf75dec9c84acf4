"""Beyond the Hock-Schittkowski set: 25 small models of the reference's test/CUTE_examples directory that its own scripts do not
run (test/runhs.sh lists hs* only), copied as data fixtures to tests/golden/cute_nl.  CPU: the `.nl` reader against central
differences, the C oracle of Algorithm::Optimize (oracle/oracle_sqp.c) to the optima tabulated for the CUTE set, and the numpy
mirror of the loop (the product's host logic, on the oracle twin of the backend) against that C oracle.  GPU: the
device-resident loop against the C oracle, bit for bit."""
import os

import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from oracle import oracle_py as orc
from oracle_backend import OracleQPInterface
from test_hs_suite import perturbed_starts

CUTE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cute_nl")
# optimal values as tabulated with the CUTE collection (Bongartz, Conn, Gould, Toint 1995 and the SIF files' "Solution" lines)
F_STAR = {"orthregb": 0.0, "fccu": 11.14911, "genhs28": 0.92717369, "lotschd": 2398.4158, "oslbqp": 6.25, "hs21mod": -95.96,
          "hatfldh": -24.5, "aircrftb": 0.0, "bt3": 4.09301056, "bt8": 1.0, "bt13": 0.0, "matrix2": 0.0, "zecevic2": -4.125,
          "zecevic4": 7.5575, "polak4": 0.0, "byrdsphr": -4.68330, "makela1": -1.41421356, "demymalo": -3.0, "gigomez1": -3.0,
          "mifflin1": -1.0, "mifflin2": -1.0, "maratos": -1.0, "booth": 0.0, "simbqp": 0.0,
          "hubfit": 0.0168935}  # hubfit: an `if` / comparison model (Huber fit), the non-smooth operators of the reader
NAMES = sorted(F_STAR)


@pytest.mark.parametrize("name", NAMES)
def test_reader_derivatives_against_central_differences(name):
    h = AmplNLP(os.path.join(CUTE_DIR, name + ".nl"))
    X = perturbed_starts(h, 2, 1)
    lam = np.random.default_rng(3).standard_normal((2, h.m))
    g, J, Hv = h.Eval_gradient(X), h.Eval_Jacobian(X), h.Eval_Hessian(X, lam)
    e = 1e-6
    for b in range(2):
        Jd = np.zeros((h.m, h.n)); Jd[np.asarray(h.J_row1) - 1, np.asarray(h.J_col1) - 1] = J[b]
        Hd = np.zeros((h.n, h.n))
        for k, (i, j) in enumerate(zip(h.H_row1, h.H_col1)):
            Hd[i - 1, j - 1] = Hv[b, k]; Hd[j - 1, i - 1] = Hv[b, k]
        for i in range(h.n):
            d = np.zeros((1, h.n)); d[0, i] = e
            xp, xm = X[b:b + 1] + d, X[b:b + 1] - d
            assert abs((h.Eval_f(xp)[0] - h.Eval_f(xm)[0]) / (2 * e) - g[b, i]) <= 1e-5 * max(1.0, np.abs(g[b]).max(initial=0.0))
            if h.m:
                assert np.abs((h.Eval_constraints(xp)[0] - h.Eval_constraints(xm)[0]) / (2 * e) - Jd[:, i]).max() <= 1e-5 * max(1.0, np.abs(J[b]).max(initial=0.0))
            # Lagrangian Hessian column: d/dx_i of (grad f + J' lam), the sign convention of the reader's Eval_Hessian
            Lg = lambda x: h.Eval_gradient(x)[0] + (np.bincount(np.asarray(h.J_col1) - 1, weights=h.Eval_Jacobian(x)[0] * lam[b][np.asarray(h.J_row1) - 1],
                                                                minlength=h.n) if h.m else 0.0)
            assert np.abs((Lg(xp) - Lg(xm)) / (2 * e) - Hd[:, i]).max() <= 1e-4 * max(1.0, np.abs(Hv[b]).max(initial=0.0))


@pytest.mark.parametrize("name", NAMES)
def test_sqp_oracle_reaches_the_tabulated_optimum(name):
    h = AmplNLP(os.path.join(CUTE_DIR, name + ".nl"))
    res = orc.SqpOracle(h, r.Options(iter_max=200)).solve_batch(perturbed_starts(h, 8, 3))
    assert (res["exitflag"] == 0).sum() >= 7, res["exitflag"]
    assert int(res["exitflag"][0]) == 0
    assert abs(res["obj"][0] - F_STAR[name]) <= 1e-3 * max(1.0, abs(F_STAR[name])), res["obj"][0]


@pytest.mark.parametrize("name", ["bt3", "genhs28", "lotschd", "zecevic4", "hatfldh", "byrdsphr"])
def test_numpy_mirror_equals_the_c_oracle(name):
    """One instance at a time (the mirror shares its init / hotstart decision across a batch, the C oracle does not)."""
    h = AmplNLP(os.path.join(CUTE_DIR, name + ".nl"))
    X = perturbed_starts(h, 3, 3)
    res_c = orc.SqpOracle(h, r.Options(iter_max=200)).solve_batch(X)
    for b in range(3):
        opt = r.Options(iter_max=200)
        mk = lambda info, qptype: r.QPhandler(info, qptype, opt, batch=1, backend=OracleQPInterface(info, qptype, opt, batch=1), refresh_ubA=True)
        res = BatchedSQP(h, x0=X[b:b + 1], options=opt, make_handler=mk).Optimize()
        assert int(res.exitflag[0]) == int(res_c["exitflag"][b]) and int(res.iters[0]) == int(res_c["iters"][b])
        assert int(res.qp_iter[0]) == int(res_c["qp_iter"][b])
        assert np.array_equal(res.x[0], res_c["x"][b])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bt3", "bt13", "fccu", "genhs28", "hatfldh", "lotschd", "matrix2", "orthregb", "oslbqp", "zecevic4", "byrdsphr", "maratos"])
def test_device_loop_equals_the_c_oracle(gpu_lib, name):
    """The device-resident loop (per-instance backend state machines) against one independent CPU solve per instance on models
    evaluated with + - * and squares only, where the NVRTC and the gcc evaluators agree bitwise: identical exit flags, outer and QP
    iteration counts and iterates."""
    from restartsqp_b200.nl_reader import DeviceNLP
    from restartsqp_b200.sqp_device import DeviceBatchedSQP
    host = AmplNLP(os.path.join(CUTE_DIR, name + ".nl"))
    dev = DeviceNLP(host)
    X = perturbed_starts(host, 64, 4)
    alg = DeviceBatchedSQP(dev, x0=X, options=r.Options(iter_max=150))
    res_d = alg.Optimize()
    res_c = orc.SqpOracle(host, r.Options(iter_max=150)).solve_batch(X)
    assert (res_d.exitflag == res_c["exitflag"]).all(), (res_d.exitflag, res_c["exitflag"])
    assert (res_d.iters == res_c["iters"]).all() and (res_d.qp_iter == res_c["qp_iter"]).all()
    fin = np.isfinite(res_c["x"]).all(axis=1)
    assert np.array_equal(res_d.x[fin], res_c["x"][fin]) and np.array_equal(res_d.obj[fin], res_c["obj"][fin])
    assert (res_d.exitflag == 0).mean() >= 0.8
    alg.close(); dev.close()
