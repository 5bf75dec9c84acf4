"""Shared problem generators and fixture loaders for the tests (test infrastructure)."""
import json
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_qp_fixtures():
    with open(os.path.join(GOLDEN, "qp_fixtures.json")) as f:
        return json.load(f)["qps"]


def load_l0_golden():
    with open(os.path.join(GOLDEN, "l0_golden.json")) as f:
        return json.load(f)["cases"]


def is_symmetric_fixture(q):
    nV = q["nV"]
    H = sp.csc_matrix((q["H_val"], q["H_rowidx"], q["H_colptr"]), shape=(nV, nV)).toarray()
    return np.abs(H - H.T).max() == 0.0


def csc(M):
    M = sp.csc_matrix(M)
    M.sort_indices()
    return M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)


def random_l1_qp(rng, n, m, convex=True, rho=None, dens=1.0):
    """A random l1-penalty trust-region QP in the layout of src/QPhandler.cpp:150-156:
    x=[p;u;v], A=[J I -I], H=[Hn 0;0 0], g=[grad; rho*1], box trust region, slacks >= 0."""
    nV, nC = n + 2 * m, m
    J = rng.standard_normal((m, n)) * (rng.random((m, n)) < dens)
    for i in range(m):
        if not J[i].any():
            J[i, rng.integers(0, n)] = rng.standard_normal()
    A = np.hstack([J, np.eye(m), -np.eye(m)]) if m else np.zeros((0, nV))
    S = rng.standard_normal((n, n))
    Hn = S @ S.T + 0.5 * np.eye(n) if convex else (S + S.T)
    H = np.zeros((nV, nV))
    H[:n, :n] = Hn
    rho = 10.0 ** rng.integers(0, 3) if rho is None else rho
    g = np.concatenate([rng.standard_normal(n), np.full(2 * m, float(rho))])
    delta = 1.0
    lb = np.concatenate([np.full(n, -delta), np.zeros(2 * m)])
    ub = np.concatenate([np.full(n, delta), np.full(2 * m, 1e18)])
    ck = rng.standard_normal(m)
    lbA, ubA = -ck.copy(), -ck.copy()
    for i in range(m):
        t = rng.integers(0, 3)
        if t == 1:
            lbA[i] = -1e18
        if t == 2:
            ubA[i] = lbA[i] + abs(rng.standard_normal())
    return dict(nV=nV, nC=nC, n=n, m=m, H=H, A=A, g=g, lb=lb, ub=ub, lbA=lbA, ubA=ubA)


def oracle_solve(orc, p, is_lp=False, max_iter=1000, Acsc=None, Hcsc=None):
    """Solve one QP with the CPU oracle; returns dict(x,y,obj,iters,status,wb,wc)."""
    Acsc = csc(p["A"]) if Acsc is None else Acsc
    Hcsc = (None if is_lp else csc(p["H"])) if Hcsc is None else Hcsc
    s = orc.OracleQP(p["nV"], p["nC"], max_iter=max_iter)
    st = s.init(Hcsc, p["g"], Acsc, p["lb"], p["ub"], p["lbA"], p["ubA"], is_lp=is_lp)
    x, y, obj, it = s.solution()
    wb, wc = s.working_set()
    return dict(x=x, y=y, obj=obj, iters=it, status=st, wb=wb, wc=wc, solver=s)
