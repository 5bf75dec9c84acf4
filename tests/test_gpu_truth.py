"""GPU output against ground truth that is NOT the co-designed CPU oracle ("parity first").

Nothing in this file imports oracle/.  The CUDA path (through the C ABI) is compared directly with
  * exhaustive working-set enumeration (numpy KKT solves) on small strictly convex l1-penalty QPs: the unique solution
    and, under strict complementarity, the unique active set -- which is therefore also what qpOASES returns;
  * HiGHS (scipy.optimize.linprog) on the penalty-steering LPs;
  * the KKT conditions evaluated in extended precision (numpy longdouble) from the returned (x, y, working set) at
    1e-10 ABSOLUTE on every strictly convex configuration, the config-4 shapes included.
    For a strictly convex QP the KKT conditions are sufficient, so passing them proves optimality without any solver;
  * QPhandler::get_active_set restated independently from src/QPhandler.cpp:600-655 (quirk 3 included).
"""
import itertools

import numpy as np
import pytest
from scipy.optimize import linprog

import restartsqp_b200 as r
from restartsqp_b200 import capi
import helpers as H

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------- ground truth
def enumerate_qp(p):
    """Unique minimiser of a strictly convex (on the x block) l1-penalty QP by enumeration of every working set.
    Returns (obj, x, y_bounds, y_constr)."""
    Hm, g, A, lb, ub, lbA, ubA = p["H"], p["g"], p["A"], p["lb"], p["ub"], p["lbA"], p["ubA"]
    nV, nC, best = p["nV"], p["nC"], None
    for sb in itertools.product((0, -1, 1), repeat=nV):
        fx = [i for i in range(nV) if sb[i] != 0]
        fr = [i for i in range(nV) if sb[i] == 0]
        x0 = np.zeros(nV)
        for i in fx:
            x0[i] = lb[i] if sb[i] < 0 else ub[i]
        if np.any(np.abs(x0) > 1e17):
            continue
        for sc in itertools.product((0, -1, 1), repeat=nC):
            ac = [i for i in range(nC) if sc[i] != 0]
            if len(ac) > len(fr):
                continue
            nf, na = len(fr), len(ac)
            K, rhs = np.zeros((nf + na, nf + na)), np.zeros(nf + na)
            K[:nf, :nf] = Hm[np.ix_(fr, fr)]
            if na:
                Aa = A[np.ix_(ac, fr)]
                K[:nf, nf:], K[nf:, :nf] = -Aa.T, Aa
                bb = np.array([lbA[i] if sc[i] < 0 else ubA[i] for i in ac])
                if np.any(np.abs(bb) > 1e17):
                    continue
                rhs[nf:] = bb - A[np.ix_(ac, fx)] @ x0[fx]
            rhs[:nf] = -(g[fr] + Hm[np.ix_(fr, fx)] @ x0[fx])
            if nf + na:
                if np.linalg.matrix_rank(K) < nf + na:
                    continue
                sol = np.linalg.solve(K, rhs)
            else:
                sol = np.zeros(0)
            x = x0.copy()
            x[fr] = sol[:nf]
            yc = np.zeros(nC)
            for k, i in enumerate(ac):
                yc[i] = sol[nf + k]
            yb = Hm @ x + g - A.T @ yc
            tol = 1e-9
            Ax = A @ x
            if np.any(x < lb - tol) or np.any(x > ub + tol) or np.any(Ax < lbA - tol) or np.any(Ax > ubA + tol):
                continue
            ok = all((sb[i] == 0 and abs(yb[i]) < 1e-7) or (sb[i] < 0 and yb[i] > -tol) or (sb[i] > 0 and yb[i] < tol)
                     for i in range(nV))
            ok = ok and all(sc[i] == 0 or (sc[i] < 0 and yc[i] > -tol) or (sc[i] > 0 and yc[i] < tol) for i in range(nC))
            if not ok:
                continue
            obj = 0.5 * x @ Hm @ x + g @ x
            if best is None or obj < best[0] - 1e-12:
                best = (obj, x, yb, yc)
    return best


def kkt_longdouble(p, x, y, wb, wc):
    """max-norm KKT residuals in extended precision from the solver's output and its raw working set
    (+1 upper, -1 lower, 0 inactive): stationarity, primal feasibility, dual sign, complementarity."""
    L = np.longdouble
    Hm, A = p["H"].astype(L), p["A"].astype(L)
    g, lb, ub, lbA, ubA = (p[k].astype(L) for k in ("g", "lb", "ub", "lbA", "ubA"))
    nV = p["nV"]
    x, yb, yc = x.astype(L), y[:nV].astype(L), y[nV:].astype(L)
    Ax = A @ x
    stat = np.abs(Hm @ x + g - A.T @ yc - yb).max()
    prim = max(0.0, float((lb - x).max()), float((x - ub).max()), float((lbA - Ax).max(initial=0.0)), float((Ax - ubA).max(initial=0.0)))
    dual = 0.0
    for yy, w in ((yb, wb), (yc, wc)):
        for i in range(len(w)):
            if w[i] == 0:
                dual = max(dual, abs(float(yy[i])))
            elif w[i] < 0:
                dual = max(dual, max(0.0, -float(yy[i])))
            else:
                dual = max(dual, max(0.0, float(yy[i])))
    comp = 0.0
    for i in range(nV):
        if wb[i] < 0:
            comp = max(comp, abs(float(yb[i] * (x[i] - lb[i]))))
        elif wb[i] > 0:
            comp = max(comp, abs(float(yb[i] * (ub[i] - x[i]))))
    for i in range(p["nC"]):
        if wc[i] < 0:
            comp = max(comp, abs(float(yc[i] * (Ax[i] - lbA[i]))))
        elif wc[i] > 0:
            comp = max(comp, abs(float(yc[i] * (ubA[i] - Ax[i]))))
    return dict(stat=float(stat), prim=prim, dual=dual, comp=comp)


def solve_dense_batch(ps, qptype=r.QPType.QP, team_size=0):
    """Instances of one shape with dense J / H blocks share the CSC pattern: one handle, per-instance values."""
    p0 = ps[0]
    nV, nC, B = p0["nV"], p0["nC"], len(ps)
    pat_A = (np.abs(np.array([p["A"] for p in ps])).sum(axis=0) != 0).astype(float)
    pat_H = (np.abs(np.array([p["H"] for p in ps])).sum(axis=0) != 0).astype(float)
    Ac, Hc = H.csc(pat_A), H.csc(pat_H)
    def ent(k, c):  # [B][nnz] values in CSC order of the shared pattern
        cols = np.repeat(np.arange(nV), np.diff(c[0]))
        return np.ascontiguousarray(np.array([q[k][c[1], cols] for q in ps]))
    s = r.CudaQPInterface(nV=nV, nC=nC, qptype=qptype, batch=B, team_size=team_size)
    s.set_csc(capi.MAT_A, Ac[0], Ac[1], ent("A", Ac))
    if qptype == r.QPType.QP:
        s.set_csc(capi.MAT_H, Hc[0], Hc[1], ent("H", Hc))
    stack = lambda k: np.ascontiguousarray(np.array([p[k] for p in ps]))
    s.set_g(stack("g")); s.set_lb(stack("lb")); s.set_ub(stack("ub"))
    if nC:
        s.set_lbA(stack("lbA")); s.set_ubA(stack("ubA"))
    s._solve(qptype, None, None, 0)
    return s


def outputs(s):
    x = s.get_optimal_solution()
    y = np.concatenate([s.get_multipliers_bounds(), s.get_multipliers_constr()], axis=1)
    wc, wb = s.get_working_set(translated=False)
    return x, y, wb, wc, s.get_obj_value(), s.get_status()


# ---------------------------------------------------------------------------------------------- tests
@pytest.mark.parametrize("shape", [(1, 0), (2, 0), (3, 0), (1, 1), (2, 1), (3, 1), (4, 1), (2, 2), (3, 2)])
def test_gpu_vs_exhaustive_enumeration(gpu_lib, shape):
    """nV + nC <= 9: the GPU's x / y / objective equal the enumerated unique solution to 1e-8, and its final working set
    equals the enumerated active set wherever strict complementarity holds."""
    n, m = shape
    rng = np.random.default_rng(9000 + 17 * n + m)
    ps = [H.random_l1_qp(rng, n, m, convex=True) for _ in range(12)]
    s = solve_dense_batch(ps)
    x, y, wb, wc, obj, st = outputs(s)
    s.close()
    nV = ps[0]["nV"]
    n_ws = 0
    for b, p in enumerate(ps):
        t = enumerate_qp(p)
        assert st[b] == 20
        assert np.abs(x[b] - t[1]).max() <= 1e-8 * max(1.0, np.abs(t[1]).max())
        assert np.abs(y[b, :nV] - t[2]).max() <= 1e-7 * max(1.0, np.abs(t[2]).max())
        if m:
            assert np.abs(y[b, nV:] - t[3]).max() <= 1e-7 * max(1.0, np.abs(t[3]).max())
        assert abs(obj[b] - t[0]) <= 1e-8 * max(1.0, abs(t[0]))
        # active set of the enumerated solution; compare where it is strictly complementary (unique active set)
        Ax = p["A"] @ t[1]
        tb = np.where(np.abs(t[1] - p["lb"]) < 1e-9, -1, np.where(np.abs(t[1] - p["ub"]) < 1e-9, 1, 0))
        tc = np.where(np.abs(Ax - p["lbA"]) < 1e-9, -1, np.where(np.abs(Ax - p["ubA"]) < 1e-9, 1, 0))
        strict = (np.abs(t[2][tb != 0]) > 1e-6).all() and (np.abs(t[3][tc != 0]) > 1e-6).all()
        if strict:
            n_ws += 1
            eq = np.abs(p["lbA"] - p["ubA"]) < 1e-9  # an active equality row may sit in the working set on either side
            assert (wb[b] == tb).all(), (wb[b], tb)
            assert (wc[b][~eq] == tc[~eq]).all() and ((wc[b][eq] != 0) == (tc[eq] != 0)).all(), (wc[b], tc)
    assert n_ws >= 8  # most random instances are non-degenerate: the working-set comparison is not vacuous


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (4, 2), (5, 5), (7, 4), (8, 6)])
def test_gpu_lp_vs_highs(gpu_lib, shape):
    n, m = shape
    rng = np.random.default_rng(7000 + 13 * n + m)
    ps = []
    for _ in range(16):
        p = H.random_l1_qp(rng, n, m, rho=1.0)
        p["g"][:n] = 0.0  # the penalty-steering LP: minimise rho*e'(u+v) (src/Algorithm.cpp:700-704)
        ps.append(p)
    s = solve_dense_batch(ps, qptype=r.QPType.LP)
    x, y, wb, wc, obj, st = outputs(s)
    s.close()
    for b, p in enumerate(ps):
        A, lbA, ubA, lb, ub, g = p["A"], p["lbA"], p["ubA"], p["lb"], p["ub"], p["g"]
        Aub, bub = [], []
        for i in range(m):
            if ubA[i] < 1e17:
                Aub.append(A[i]); bub.append(ubA[i])
            if lbA[i] > -1e17:
                Aub.append(-A[i]); bub.append(-lbA[i])
        ref = linprog(g, A_ub=np.array(Aub), b_ub=np.array(bub),
                      bounds=[(lb[i], None if ub[i] > 1e17 else ub[i]) for i in range(p["nV"])], method="highs")
        assert st[b] == 20
        assert abs(g @ x[b] - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun))
        Ax = A @ x[b]
        assert max(0.0, (lbA - Ax).max(), (Ax - ubA).max(), (lb - x[b]).max(), (x[b] - ub).max()) < 1e-9


@pytest.mark.parametrize("shape", [(3, 2), (6, 4), (10, 7), (16, 8), (23, 12), (30, 20)])
def test_gpu_kkt_1e10_absolute_small(gpu_lib, shape):
    """north_star gate: KKT residuals <= 1e-10 absolute on strictly convex QPs, evaluated independently in extended precision."""
    n, m = shape
    rng = np.random.default_rng(5000 + 11 * n + m)
    ps = [H.random_l1_qp(rng, n, m, convex=True, rho=1.0) for _ in range(16)]
    s = solve_dense_batch(ps)
    x, y, wb, wc, obj, st = outputs(s)
    s.close()
    for b, p in enumerate(ps):
        assert st[b] == 20
        k = kkt_longdouble(p, x[b], y[b], wb[b], wc[b])
        assert max(k.values()) <= 1e-10, k  # absolute


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024])
def test_gpu_kkt_1e10_absolute_config4(gpu_lib, n):
    """BASELINE configs[3] shapes (synthetic sparse QP, m = n/2, 1 % density; from n = 64 on the cluster kernel: TMA-staged DMMA
    refactorisation shared by up to 8 CTAs per QP).  At n = 512 and 1024 the scalar CPU oracle needs minutes per instance, so the
    check is the solver-independent one: the KKT conditions in extended precision, which for a strictly convex QP prove optimality."""
    B = 4 if n <= 512 else 2
    d = H.synthetic_large_qp(n, batch=B)
    s = r.CudaQPInterface(nV=d["nV"], nC=d["nC"], qptype=r.QPType.QP, batch=B, keep_state=False,
                          options=r.Options(qp_maxiter=6 * n))
    s.set_csc(capi.MAT_A, *d["Ac"]); s.set_csc(capi.MAT_H, *d["Hc"])
    s.set_g(d["g"]); s.set_lb(d["lb"]); s.set_ub(d["ub"]); s.set_lbA(d["lbA"]); s.set_ubA(d["ubA"])
    s._solve(r.QPType.QP, None, None, 0)
    x, y, wb, wc, obj, st = outputs(s)
    s.close()
    import scipy.sparse as sp
    A = sp.csc_matrix((d["Ac"][2], d["Ac"][1], d["Ac"][0]), shape=(d["nC"], d["nV"])).toarray()
    Hm = sp.csc_matrix((d["Hc"][2], d["Hc"][1], d["Hc"][0]), shape=(d["nV"], d["nV"])).toarray()
    for b in range(B):
        assert st[b] == 20
        p = dict(nV=d["nV"], nC=d["nC"], H=Hm, A=A, g=d["g"][b], lb=d["lb"][b], ub=d["ub"][b], lbA=d["lbA"][b], ubA=d["ubA"][b])
        k = kkt_longdouble(p, x[b], y[b], wb[b], wc[b])
        assert max(k.values()) <= 1e-10, k  # absolute


def test_get_active_set_reference_formula(gpu_lib):
    """QPhandler::get_active_set (src/QPhandler.cpp:600-655): geometric active set at tolerance sqrt_m_eps; the non-QORE
    branch reads ubA for BOTH constraint sides (quirk 3, :641-642).  Restated here line by line, independently."""
    rng = np.random.default_rng(31)
    n, m, B = 5, 3, 32
    base = H.random_l1_qp(rng, n, m, convex=True)
    info = r.NLPInfo(nVar=n, nCon=m)
    qh = r.QPhandler(info, r.QPType.QP, batch=B)
    si = qh.solverInterface_
    Ac, Hc = H.csc(base["A"]), H.csc(base["H"])
    si.set_csc(capi.MAT_A, *Ac); si.set_csc(capi.MAT_H, *Hc)
    g = np.tile(base["g"], (B, 1)); g[:, :n] += rng.standard_normal((B, n))
    lbA = np.tile(base["lbA"], (B, 1)); ubA = np.tile(base["ubA"], (B, 1))
    si.set_g(g); si.set_lb(np.tile(base["lb"], (B, 1))); si.set_ub(np.tile(base["ub"], (B, 1))); si.set_lbA(lbA); si.set_ubA(ubA)
    qh.solveQP()
    A_c, A_b = qh.get_active_set()
    x = qh.get_optimal_solution()
    eps = 1.0e-8  # sqrt_m_eps, include/sqphot/Utils.hpp
    seen = set()
    for b in range(B):
        Ax = base["A"] @ x[b]
        for i in range(n + 2 * m):
            lo, hi = abs(x[b, i] - base["lb"][i]) < eps, abs(base["ub"][i] - x[b, i]) < eps
            want = -99 if (lo and hi) else (-1 if lo else (1 if hi else 0))
            assert A_b[b, i] == want
            seen.add(want)
        for i in range(m):
            lo, hi = abs(Ax[i] - ubA[b, i]) < eps, abs(ubA[b, i] - Ax[i]) < eps  # both sides read ubA (quirk 3)
            want = -99 if (lo and hi) else (-1 if lo else (1 if hi else 0))
            if abs(abs(Ax[i] - ubA[b, i]) - eps) > 1e-12:  # A x is recomputed on the device: skip the knife edge
                assert A_c[b, i] == want
    assert {-1, 0}.issubset(seen)
    si.close()
