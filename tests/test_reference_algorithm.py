"""PINS the oracle of the SQP loop (oracle/oracle_sqp.c) -- and through it the device-resident loop, which the GPU tests compare
with that oracle bit for bit -- against the reference's REAL code: src/Algorithm.cpp (initialization + Optimize), src/SQPTNLP.cpp
and src/QPhandler.cpp, unmodified except for integration/restartsqp_cuda_backend.patch, compiled from a scratch copy into
oracle/_ref/algorithm_nl* (oracle/algorithm_test.cpp; dev container only) and run one instance at a time the way
test/simple_test.cpp runs them.  Ipopt / ASL are absent: the NLP comes in as an Ipopt::TNLP over the C evaluator the `.nl` reader
generates (the arithmetic oracle_sqp.c evaluates).  qpOASES / QORE are absent: the QP / LP backend is the QORE-layout CUDA plugin
(CudaQOREInterface through QPhandler's QORE branches), linked here with the CPU twin of the C ABI over the oracle
(oracle/capi_twin.cpp), on the GPU box with libsqpb200.so.

What must hold: the reference's loop and the restated loop make the same decisions on the same QP solutions -- identical exit
flag, outer and QP iteration counts, final iterate and objective, bit for bit.  What differs, and why, is asserted too:
  * the QORE setters clip bounds to +-1e18 (include/sqphot/QOREInterface.hpp:142-155); oracle_sqp.c and the batched kernels take
    c_u - c_k = 1e19 as it is.  The `_noclip` build removes that difference; with it (the plugin as shipped) non-convex models with
    one-sided constraints take other QP iterations;
  * labels of failures: QORE's get_status maps every state but four to QPERROR_UNKNOWN (src/QOREInterface.cpp:425-438); a QP that
    is solved but fails the KKT test leaves the reference with exitflag QP_OPTIMAL (src/Algorithm.cpp:68-71), which the oracle
    labels QPERROR_INTERNAL_ERROR;
  * a NaN KKT error passes the reference's `KKT_error > tol` test (the NaN QP is accepted, the next setupQP throws QP_UNCHANGED
    uncaught); the oracle and the library reject it;
  * (outside the models tested here: hs099, hs99exp, hs107) QORE's KKT tolerance is 1e-5, the oracle's 1e-6.
profiles/r2_reference_loop_pin.md holds the same comparison over all 149 fixture models."""
import glob
import os
import subprocess

import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, write_model_file
from oracle import oracle_py as orc
from test_hs_suite import HS_DIR, perturbed_starts
from test_cute_suite import CUTE_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
TWIN, NOCLIP, PRODUCT = (os.path.join(REF, n) for n in ("algorithm_nl_twin", "algorithm_nl_twin_noclip", "algorithm_nl"))
needs_ref = pytest.mark.skipif(not (os.path.exists(TWIN) and os.path.exists(NOCLIP)),
                               reason="oracle/_ref/algorithm_nl_twin not built (needs /root/reference at build time)")
HS = ["hs071", "hs043", "hs015", "hs113", "hs083", "hs093", "hs106", "hs108", "hs116", "hs118", "hs100", "hs035", "hs024", "hs012", "hs076"]
CUTE = ["bt3", "lotschd", "genhs28", "fccu", "zecevic4", "hatfldh", "byrdsphr", "orthregb"]
B = 6


def run_all(binary, name, tmp_path, mode="qore", timeout=300, count=B):
    d = HS_DIR if name.startswith("hs") else CUTE_DIR
    h = AmplNLP(os.path.join(d, name + ".nl"))
    X = perturbed_starts(h, B, 4)
    res = orc.SqpOracle(h, r.Options()).solve_batch(X)  # also compiles the C evaluator into oracle/_gen
    model = str(tmp_path / (name + ".model"))
    write_model_file(h, model, X)
    ev = sorted(glob.glob(os.path.join(ROOT, "oracle", "_gen", "nlp_%s_*.so" % name)), key=os.path.getmtime)[-1]
    out = []
    for k in range(count):
        p = subprocess.run([binary, model, ev, str(k)] + ([mode] if mode else []), capture_output=True, text=True, timeout=timeout)
        t = p.stdout.split()
        if p.returncode != 0:
            out.append(dict(rc=p.returncode, msg=p.stdout.strip()))
        else:
            out.append(dict(rc=0, exitflag=int(t[0]), iters=int(t[1]), qp_iter=int(t[2]), obj=float.fromhex(t[3]),
                            x=np.array([float.fromhex(v) for v in t[4:]])))
    return res, out


@needs_ref
@pytest.mark.parametrize("name", HS + CUTE)
def test_reference_optimize_equals_the_oracle_of_the_loop(name, tmp_path):
    res, out = run_all(NOCLIP, name, tmp_path)
    for k, o in enumerate(out):
        ex_o = int(res["exitflag"][k])
        if o["rc"] != 0:  # the only way out of the reference with an exception: QP_UNCHANGED after a NaN QP was accepted
            assert "QP is not changed" in o["msg"] and ex_o == 21 and not np.isfinite(res["KKT_error"][k]) or ex_o in (7, 21), (name, k, o)
            continue
        assert o["iters"] == int(res["iters"][k]) and o["qp_iter"] == int(res["qp_iter"][k]), (name, k)
        assert np.array_equal(o["x"], res["x"][k]) and o["obj"] == res["obj"][k], (name, k)
        if o["exitflag"] != ex_o:  # failure labels only (see the module docstring)
            assert (o["exitflag"], ex_o) in ((30, 28), (30, 21), (30, 26), (30, 27), (30, 29), (20, 21)), (name, k, o["exitflag"], ex_o)


@needs_ref
@pytest.mark.parametrize("name", ["hs071", "hs043", "hs113", "hs100", "hs118", "bt3", "lotschd", "fccu"])
def test_reference_optimize_with_the_plugin_as_shipped(name, tmp_path):
    """The same with the QORE setters' clipping in place (the plugin a maintainer links): identical wherever no constraint bound is
    infinite or the QPs are convex enough for the value of an inactive far bound not to matter."""
    res, out = run_all(TWIN, name, tmp_path)
    for k, o in enumerate(out):
        assert o["rc"] == 0 and o["exitflag"] == int(res["exitflag"][k]) and o["iters"] == int(res["iters"][k]), (name, k)
        assert o["qp_iter"] == int(res["qp_iter"][k]) and np.array_equal(o["x"], res["x"][k]) and o["obj"] == res["obj"][k], (name, k)


@needs_ref
def test_reference_second_order_correction_equals_the_oracle(tmp_path):
    """The opt-in second-order correction (src/Algorithm.cpp:1140-1211) of the reference's real code against the oracle's: same exit
    flags, outer and QP iteration counts and iterates, bit for bit, on 7 models x 6 starts (on five of them the correction changes
    the run).  Running the reference's code here is what found the one deviation the restatements had: update_radius takes the
    norm of the CORRECTED step after an accepted correction (p_k_->getInfNorm(), src/Algorithm.cpp:822)."""
    exact = total = 0
    for name in ["hs006", "hs043", "hs100", "hs015", "hs113", "hs071", "hs038"]:
        h = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
        X = perturbed_starts(h, B, 4)
        res = orc.SqpOracle(h, r.Options(second_order_correction=True)).solve_batch(X)
        model = str(tmp_path / (name + ".model"))
        write_model_file(h, model, X)
        ev = sorted(glob.glob(os.path.join(ROOT, "oracle", "_gen", "nlp_%s_*.so" % name)), key=os.path.getmtime)[-1]
        for k in range(B):
            p = subprocess.run([NOCLIP, model, ev, str(k), "qore", "soc"], capture_output=True, text=True, timeout=300)
            assert p.returncode == 0, (name, k, p.stdout)
            t = p.stdout.split()
            x = np.array([float.fromhex(v) for v in t[4:]])
            assert (int(t[0]), int(t[1]), int(t[2])) == (int(res["exitflag"][k]), int(res["iters"][k]), int(res["qp_iter"][k])), (name, k)
            assert np.abs(x - res["x"][k]).max() <= 1e-14 * max(1.0, np.abs(x).max()), (name, k)
            exact += int(np.array_equal(x, res["x"][k]))
            total += 1
    assert total == 42 and exact == 42


@needs_ref
def test_clipped_far_bounds_change_qp_iterations_not_results(tmp_path):
    """hs015 (non-convex, one-sided constraints): with the bounds clipped to 1e18 two of six runs take other QP iterations
    (flipped bounds land on the far bound's value) and arrive at the same iterates."""
    res, out = run_all(TWIN, "hs015", tmp_path)
    assert all(o["rc"] == 0 and np.array_equal(o["x"], res["x"][k]) and o["iters"] == int(res["iters"][k]) for k, o in enumerate(out))
    assert any(o["qp_iter"] != int(res["qp_iter"][k]) for k, o in enumerate(out))


@needs_ref
def test_qpoases_layout_branch_of_the_reference_stops_at_its_stale_ubA(tmp_path):
    """Through QPhandler's non-QORE branches (Solver CUDA_B200, the qpOASES data layout) the reference never refreshes ubA after the
    first iteration (src/QPhandler.cpp:358-360): on HS071 the second QP has lbA > ubA for the equality
    constraint, its solution fails QPhandler's KKT test and Optimize ends after one iteration with exitflag QP_OPTIMAL (20).  This
    is the reference's behaviour with any backend of that layout; the batched drivers refresh both sides, as its QORE branch does."""
    res, out = run_all(TWIN, "hs071", tmp_path, mode="")
    assert all(o["rc"] == 0 and o["exitflag"] != 0 for o in out)  # no run reaches OPTIMAL ...
    assert sum(o["exitflag"] == 20 and o["iters"] <= 2 for o in out) >= 4  # ... most stop at the second QP
    assert (res["exitflag"] == 0).all()  # where the loop with both sides refreshed converges from every start


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first GPU run of this driver happens at round end (the builder's GPU budget was spent); "
                                        "the same objects pass on the CPU twin")
@pytest.mark.parametrize("name", ["hs071", "hs043", "bt3"])
def test_reference_optimize_on_the_gpu_backend(gpu_lib, name, tmp_path):
    if not (os.path.exists(PRODUCT) and os.path.exists(TWIN)):
        pytest.skip("oracle/_ref/algorithm_nl not built (needs /root/reference at build time)")
    res, out = run_all(PRODUCT, name, tmp_path, timeout=30, count=3)  # bounded: this program has not run on a GPU before
    for k, o in enumerate(out):
        assert o["rc"] == 0 and o["exitflag"] == int(res["exitflag"][k]) and o["iters"] == int(res["iters"][k]), (name, k)
        assert o["qp_iter"] == int(res["qp_iter"][k]) and np.array_equal(o["x"], res["x"][k]) and o["obj"] == res["obj"][k], (name, k)
