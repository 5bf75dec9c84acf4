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
NEEDS_FUNCS = {"hs068", "hs069"}  # imported AMPL function `myerf` (F segment) = the standard normal distribution function


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


def test_chunked_cuda_source_equals_one_piece_program(tmp_path):
    """AmplNLP.cuda_source cuts long straight-line programs into __noinline__ device functions (large DAGs: hs025, hs088-092,
    hs105).  The chunked source, compiled here as plain C++ with the CUDA keywords defined away, must reproduce the numpy
    evaluator generated from the same DAG at every output (same operations in the same order)."""
    import ctypes, os, subprocess
    from restartsqp_b200.nl_reader import AmplNLP
    nlp = AmplNLP(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hs_nl", "hs085.nl"))
    src = nlp.cuda_source(piece=200)
    assert "nlp_piece_all_3" in src and "nl_exp" in src
    n, m, zJ, zH = nlp.n, nlp.m, len(nlp.jac_nodes), len(nlp._hess_terms)
    harness = """
#include <math.h>
#include <stddef.h>
#define __device__
#define __noinline__
#define __forceinline__ inline
#define __global__
#define __restrict__
struct Idx { int x; };
static Idx blockIdx = {0}, blockDim = {1}, threadIdx = {0};
%s
extern "C" void run_all(const double* x, const double* lam, double* f, double* c, double* g, double* j, double* h) { nlp_eval_all(1, x, lam, f, c, g, j, h); }
extern "C" void run_fc(const double* x, double* f, double* c) { nlp_eval_fc(1, x, f, c); }
""" % src.replace('extern "C" __global__', 'static')
    cpp, so = tmp_path / "chunked.cpp", tmp_path / "chunked.so"
    cpp.write_text(harness)
    subprocess.run(["g++", "-O0", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(so), str(cpp)], check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(85)
    x0, _ = nlp.Get_starting_point()
    dp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for _ in range(3):
        x = x0 * (1 + 0.01 * rng.standard_normal(n))
        lam = rng.standard_normal(m)
        f, c, g, j, h = np.zeros(1), np.zeros(m), np.zeros(n), np.zeros(zJ), np.zeros(zH)
        lib.run_all(dp(x), dp(lam), dp(f), dp(c), dp(g), dp(j), dp(h))
        rel = lambda a, b: float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))
        assert rel(f, nlp.Eval_f(x)) < 1e-13 and rel(c, nlp.Eval_constraints(x)[0]) < 1e-13
        assert rel(g, nlp.Eval_gradient(x)[0]) < 1e-13 and rel(j, nlp.Eval_Jacobian(x)[0]) < 1e-13
        assert rel(h, nlp.Eval_Hessian(x, lam)[0]) < 1e-12
        f2, c2 = np.zeros(1), np.zeros(m)
        lib.run_fc(dp(x), dp(f2), dp(c2))
        assert f2[0] == f[0] and np.array_equal(c2, c)


def test_imported_function_myerf_is_the_normal_distribution_function():
    """hs068 / hs069 call the imported AMPL function `myerf`; the function library is not part of the reference.  Hock-Schittkowski
    define the problems with the standard normal distribution function Phi: with it the tabulated optimum of hs068 (x* in the
    file's variable order, f* = -0.920425) is feasible to the printed digits."""
    p = AmplNLP(os.path.join(HS_DIR if "HS_DIR" in globals() else os.path.dirname(FILES[0]), "hs068.nl"))
    xs = np.array([3.64617, 0.0678587, 0.000266, 0.894862])
    assert abs(p.Eval_f(xs)[0] + 0.920425) < 1e-5
    assert np.abs(p.Eval_constraints(xs)).max() < 1e-5
