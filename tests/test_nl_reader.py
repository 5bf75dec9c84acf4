"""The .nl reader / batched NLP evaluator (restartsqp_b200/nl_reader.py) on the reference's own
Hock-Schittkowski files (tests/golden/hs_nl = test/CUTE_examples/hs*.nl): structure and values against the hand-written HS071
of the driver, first and second derivatives against central differences on every file."""
import glob
import os

import numpy as np
import pytest

from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import HS071

HS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hs_nl")
FILES = sorted(glob.glob(os.path.join(HS_DIR, "hs*.nl")))
NEEDS_FUNCS = {"hs068", "hs069"}  # imported AMPL functions (F segments): not readable without amplfunc libraries


def test_all_hs_files_present():
    assert len(FILES) == 124


def test_hs071_matches_hand_written_model():
    p, h = AmplNLP(os.path.join(HS_DIR, "hs071.nl")), HS071()
    info = p.Get_nlp_info()
    assert (info.nVar, info.nCon, info.nnz_jac_g, info.nnz_h_lag) == (4, 2, 8, 10)  # hs071.nl:2,8
    for a in ("J_row1", "J_col1", "H_row1", "H_col1"):
        assert (getattr(p, a) == getattr(h, a)).all(), a
    xl, xu, cl, cu = p.Get_bounds_info()
    assert xl.tolist() == [1.0] * 4 and xu.tolist() == [5.0] * 4 and cl.tolist() == [25.0, 40.0] and cu[1] == 40.0 and cu[0] >= 1e18
    assert p.Get_starting_point()[0].tolist() == [1.0, 5.0, 5.0, 1.0]
    rng = np.random.default_rng(0)
    x, lam = 1 + 4 * rng.random((7, 4)), rng.standard_normal((7, 2))
    for a, b in [(p.Eval_f(x), h.Eval_f(x)), (p.Eval_gradient(x), h.Eval_gradient(x)), (p.Eval_constraints(x), h.Eval_constraints(x)),
                 (p.Eval_Jacobian(x), h.Eval_Jacobian(x)), (p.Eval_Hessian(x, lam), h.Eval_Hessian(x, lam))]:
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-3] for f in FILES])
def test_derivatives_against_central_differences(path):
    name = os.path.basename(path)[:-3]
    if name in NEEDS_FUNCS:
        with pytest.raises(NotImplementedError):
            AmplNLP(path)
        return
    p = AmplNLP(path)
    n, m = p.n, p.m
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x0, _ = p.Get_starting_point()
    xl, xu, _, _ = p.Get_bounds_info()
    B, e = 3, 1e-6
    x = np.clip(x0, xl, xu)[None, :] + 1e-3 * rng.random((B, n))
    lam = rng.standard_normal((B, m))
    g, J, Hh = p.Eval_gradient(x), p.Eval_Jacobian(x), p.Eval_Hessian(x, lam)
    assert np.isfinite(g).all() and np.isfinite(J).all() and np.isfinite(Hh).all()
    # triplets: 1-based, Jacobian column-major, Hessian upper triangle by columns, no duplicates
    jk = list(zip(p.J_col1.tolist(), p.J_row1.tolist())); hk = list(zip(p.H_col1.tolist(), p.H_row1.tolist()))
    assert jk == sorted(set(jk)) and hk == sorted(set(hk)) and all(r_ <= c_ for c_, r_ in hk)

    def scatter_J(v):
        out = np.zeros((B, m, n))
        for k in range(len(p.J_row1)):
            out[:, p.J_row1[k] - 1, p.J_col1[k] - 1] = v[:, k]
        return out

    def lag_grad(xx):
        return p.Eval_gradient(xx) + np.einsum("bm,bmn->bn", lam, scatter_J(p.Eval_Jacobian(xx)))

    gfd, Jfd, Hfd = np.zeros((B, n)), np.zeros((B, m, n)), np.zeros((B, n, n))
    for j in range(n):
        xp, xm = x.copy(), x.copy()
        xp[:, j] += e; xm[:, j] -= e
        gfd[:, j] = (p.Eval_f(xp) - p.Eval_f(xm)) / (2 * e)
        if m:
            Jfd[:, :, j] = (p.Eval_constraints(xp) - p.Eval_constraints(xm)) / (2 * e)
        Hfd[:, :, j] = (lag_grad(xp) - lag_grad(xm)) / (2 * e)
    Hfull = np.zeros((B, n, n))
    for k in range(len(p.H_row1)):
        i, j = p.H_row1[k] - 1, p.H_col1[k] - 1
        Hfull[:, i, j] = Hh[:, k]
        Hfull[:, j, i] = Hh[:, k]
    sc = lambda a: max(1.0, float(np.abs(a).max())) if a.size else 1.0
    assert np.abs(g - gfd).max() <= 1e-4 * sc(gfd)
    assert np.abs(scatter_J(J) - Jfd).max() <= 1e-4 * sc(Jfd) if m else True
    assert np.abs(Hfull - Hfd).max() <= 1e-4 * sc(Hfd)


def test_cuda_source_is_emitted_for_every_output():
    p = AmplNLP(os.path.join(HS_DIR, "hs071.nl"))
    src = p.cuda_source()
    assert "nlp_eval" in src and src.count("hess[") == 10 and src.count("jac[") == 8 and src.count("grad[") == 4


def test_generated_cuda_compiles_with_nvrtc_for_sm100a():
    """sqpb200_nlp_compile needs no GPU: NVRTC cross-compiles the generated kernels (loading/launching them does)."""
    import ctypes as C
    from restartsqp_b200 import _capi as capi
    L = capi.lib()
    for name in ("hs071", "hs099"):
        p = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
        h, log = C.c_void_p(), C.create_string_buffer(4096)
        rc = L.sqpb200_nlp_compile(p.cuda_source().encode(), p.n, p.m, len(p.J_row1), len(p.H_row1), log, 4096, C.byref(h))
        assert rc == 0, L.sqpb200_nlp_last_error()
        assert L.sqpb200_nlp_cubin_size(h) > 1000
        L.sqpb200_nlp_destroy(h)
    h = C.c_void_p()
    assert L.sqpb200_nlp_compile(b"this is not CUDA", 1, 0, 0, 0, None, 0, C.byref(h)) != 0
