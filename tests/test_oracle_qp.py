"""Validate the CPU active-set oracle (oracle/oracle_qp.c) with ground truth that does not need qpOASES
(SURVEY.md section 8c): exhaustive working-set enumeration on small strictly convex QPs, HiGHS on LPs,
and the reference's own KKT acceptance test (1e-6) on the dumped QP fixtures."""
import itertools

import numpy as np
import pytest
from scipy.optimize import linprog

from oracle import oracle_py as orc
import helpers as H


def brute_force(p):
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


@pytest.mark.parametrize("seed", range(6))
def test_convex_qp_against_enumeration(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(5):
        n, m = int(rng.integers(1, 4)), int(rng.integers(0, 3))
        p = H.random_l1_qp(rng, n, m, convex=True)
        r = H.oracle_solve(orc, p)
        b = brute_force(p)
        assert r["status"] == 20
        assert np.abs(r["x"] - b[1]).max() <= 1e-8 * max(1.0, np.abs(b[1]).max())
        assert np.abs(r["y"][:p["nV"]] - b[2]).max() <= 1e-7 * max(1.0, np.abs(b[2]).max())
        if m:
            assert np.abs(r["y"][p["nV"]:] - b[3]).max() <= 1e-7 * max(1.0, np.abs(b[3]).max())
        assert abs(r["obj"] - b[0]) <= 1e-8 * max(1.0, abs(b[0]))


@pytest.mark.parametrize("seed", range(4))
def test_lp_against_highs(seed):
    rng = np.random.default_rng(200 + seed)
    for _ in range(5):
        n, m = int(rng.integers(1, 8)), int(rng.integers(1, 6))
        p = H.random_l1_qp(rng, n, m, rho=1.0)
        p["g"][:n] = 0.0  # the penalty-steering LP: minimise rho*e'(u+v) (src/Algorithm.cpp:700-704)
        r = H.oracle_solve(orc, p, is_lp=True, max_iter=100)
        A, lbA, ubA, lb, ub, g = p["A"], p["lbA"], p["ubA"], p["lb"], p["ub"], p["g"]
        Aub, bub = [], []
        for i in range(m):
            if ubA[i] < 1e17:
                Aub.append(A[i]); bub.append(ubA[i])
            if lbA[i] > -1e17:
                Aub.append(-A[i]); bub.append(-lbA[i])
        ref = linprog(g, A_ub=np.array(Aub), b_ub=np.array(bub),
                      bounds=[(lb[i], None if ub[i] > 1e17 else ub[i]) for i in range(p["nV"])], method="highs")
        assert r["status"] == 20
        assert abs(g @ r["x"] - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun))
        Ax = A @ r["x"]
        assert max(0.0, (lbA - Ax).max(), (Ax - ubA).max(), (lb - r["x"]).max(), (r["x"] - ub).max()) < 1e-9


FIX = H.load_qp_fixtures()
# Dumped QPs on which the oracle reaches a KKT point that passes the reference's own acceptance test.
# The others are the stress cases of SURVEY.md section 4.3: rho = 1e8 (absolute 1e-6 test cannot pass at
# that scale), strongly non-convex H, or a non-symmetric H array (not a QP); for those we only require
# termination with an Exitflag.
KKT_OK = {"QORE_hs015", "QORE_hs018", "QORE_hs024", "QORE_hs029", "QORE_hs034", "QORE_hs037", "QORE_hs038", "QORE_hs056",
          "QORE_hs072", "QORE_hs089", "QORE_hs091", "QORE_hs104", "QORE_hs116", "hs034_hpp", "hs039_hpp", "hs066_hpp"}


@pytest.mark.parametrize("q", FIX, ids=[q["name"] for q in FIX])
def test_dumped_qp_fixtures(q):
    nV, nC = q["nV"], q["nC"]
    A = (q["A_colptr"], q["A_rowidx"], q["A_val"])
    Hc = (q["H_colptr"], q["H_rowidx"], q["H_val"])
    s = orc.OracleQP(nV, nC)
    st = s.init(Hc, q["g"], A, q["lb"], q["ub"], q["lbA"], q["ubA"])
    assert 20 <= st <= 30
    x, y, obj, it = s.solution()
    assert np.all(np.isfinite(x)) and np.all(np.isfinite(y))
    if q["name"] in KKT_OK:
        wb, wc = s.working_set()
        Ax = orc.csc_times(nC, nV, *A, x)
        Wb, Wc = orc.translate_working_set(wb, wc, x, Ax, q["lb"], q["ub"], q["lbA"], q["ubA"])
        ok, res = orc.kkt_residuals(nV, nC, A, Hc, q["g"], q["lb"], q["ub"], q["lbA"], q["ubA"], x, y, Wb, Wc)
        assert st == 20 and ok, res


def test_hotstart_vectors_and_matrices():
    """hotstart(g,lb,ub,lbA,ubA) and hotstart(H,g,A,...) reach the same point as a cold start on the new data
    (strictly convex => unique solution)."""
    rng = np.random.default_rng(5)
    for _ in range(10):
        n, m = int(rng.integers(2, 6)), int(rng.integers(1, 4))
        p = H.random_l1_qp(rng, n, m)
        Ac, Hc = H.csc(p["A"]), H.csc(p["H"])
        r0 = H.oracle_solve(orc, p, Acsc=Ac, Hcsc=Hc)
        s = r0["solver"]
        # new vectors
        p2 = dict(p)
        p2["g"] = p["g"] + np.concatenate([0.3 * rng.standard_normal(n), np.zeros(2 * m)])
        p2["lbA"] = p["lbA"] - 0.2
        p2["ubA"] = p["ubA"] + 0.1
        assert s.hotstart(p2["g"], p2["lb"], p2["ub"], p2["lbA"], p2["ubA"]) == 20
        x_hot = s.solution()[0]
        r_cold = H.oracle_solve(orc, p2, Acsc=Ac, Hcsc=Hc)
        assert np.abs(x_hot - r_cold["x"]).max() < 1e-8
        # new matrix values on the same pattern
        Hv2 = Hc[2] * (1.0 + 0.05 * rng.random(len(Hc[2])))
        Hd = np.zeros_like(p["H"])
        Hp_, Hi_ = Hc[0], Hc[1]
        for c in range(p["nV"]):
            for e in range(Hp_[c], Hp_[c + 1]):
                Hd[Hi_[e], c] = Hv2[e]
        Hd = 0.5 * (Hd + Hd.T) + 0.1 * np.diag((np.arange(p["nV"]) < n).astype(float))
        Hc2 = H.csc(Hd)
        if not (np.array_equal(Hc2[0], Hc[0]) and np.array_equal(Hc2[1], Hc[1])):
            continue
        Av2 = Ac[2].copy()
        assert s.hotstart_matrices(Hc2[2], Av2, p2["g"], p2["lb"], p2["ub"], p2["lbA"], p2["ubA"]) == 20
        x_hot2 = s.solution()[0]
        p3 = dict(p2)
        p3["H"] = Hd
        r_cold2 = H.oracle_solve(orc, p3, Acsc=Ac, Hcsc=Hc2)
        assert np.abs(x_hot2 - r_cold2["x"]).max() < 1e-8


def test_reinit_from_previous_solution():
    """The matrix-status flip of optimizeQP (src/qpOASESInterface.cpp:202-207): init(H, g, A, ..., x_qp, y_qp, &bounds).  On a
    strictly convex QP it reaches the point a cold start reaches on the new data; with unchanged data the guess is already
    optimal (no working-set change); and the working set it ends in satisfies the KKT test."""
    rng = np.random.default_rng(11)
    done = 0
    for _ in range(12):
        n, m = int(rng.integers(2, 7)), int(rng.integers(1, 5))
        p = H.random_l1_qp(rng, n, m)
        Ac, Hc = H.csc(p["A"]), H.csc(p["H"])
        r0 = H.oracle_solve(orc, p, Acsc=Ac, Hcsc=Hc)
        if r0["status"] != 20:
            continue
        s = r0["solver"]
        args = (p["g"], p["lb"], p["ub"], p["lbA"], p["ubA"])
        assert s.reinit(Hc[2], Ac[2], *args) == 20
        x1, y1, _, it1 = s.solution()
        assert it1 == 0 and np.abs(x1 - r0["x"]).max() < 1e-10 and np.abs(y1 - r0["y"]).max() < 1e-8
        Hv2 = Hc[2] * 1.2
        Av2 = Ac[2] * (1.0 + 0.05 * rng.standard_normal(len(Ac[2])) * (np.abs(np.abs(Ac[2]) - 1.0) > 1e-12))
        g2 = p["g"] + np.concatenate([0.3 * rng.standard_normal(n), np.zeros(2 * m)])
        assert s.reinit(Hv2, Av2, g2, *args[1:]) == 20
        x2, y2, _, _ = s.solution()
        A2, H2 = (Ac[0], Ac[1], Av2), (Hc[0], Hc[1], Hv2)
        cold = H.oracle_solve(orc, dict(p, g=g2), Acsc=A2, Hcsc=H2)
        assert np.abs(x2 - cold["x"]).max() < 1e-8
        wb, wc = s.working_set()
        Ax = orc.csc_times(p["nC"], p["nV"], *A2, x2)
        Wb, Wc = orc.translate_working_set(wb, wc, x2, Ax, *args[1:])
        ok, res = orc.kkt_residuals(p["nV"], p["nC"], A2, H2, g2, *args[1:], x2, y2, Wb, Wc)
        assert ok, res
        done += 1
    assert done >= 8
