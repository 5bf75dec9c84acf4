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


def load_qore_golden():
    """Compressed-row (QORE layout) arrays of the reference's SpHbMat for the triplets of l0_golden.json, case by case."""
    with open(os.path.join(GOLDEN, "qore_golden.json")) as f:
        return json.load(f)["cases"]


def load_qore_raw_fixtures():
    """The 18 `.log` dumps unconverted (QORE layout, explicit zeros kept): inputs of the QORE data constructor."""
    with open(os.path.join(GOLDEN, "qp_fixtures_qore.json")) as f:
        return json.load(f)["qps"]


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


def oracle_solve(orc, p, is_lp=False, max_iter=1000, Acsc=None, Hcsc=None, force_error_branch=False):
    """Solve one QP with the CPU oracle; returns dict(x,y,obj,iters,status,wb,wc)."""
    Acsc = csc(p["A"]) if Acsc is None else Acsc
    Hcsc = (None if is_lp else csc(p["H"])) if Hcsc is None else Hcsc
    s = orc.OracleQP(p["nV"], p["nC"], max_iter=max_iter)
    st = s.init(Hcsc, p["g"], Acsc, p["lb"], p["ub"], p["lbA"], p["ubA"], is_lp=is_lp)
    it = s.solution()[3]
    if st != 20 or force_error_branch:  # optimizeQP -> handle_error (src/qpOASESInterface.cpp:160-162, 686-758)
        st, added = s.handle_error(force_guess=force_error_branch)
        it += added
    x, y, obj, _ = s.solution()
    wb, wc = s.working_set()
    return dict(x=x, y=y, obj=obj, iters=it, status=st, wb=wb, wc=wc, solver=s)


def synthetic_large_qp(n, seed=None, dens=0.01, batch=1):
    """n variables, m = n/2 constraints, Jacobian entries non-zero with probability `dens` (at least
    one per row), H = S + (norm1(S) + 1) I on the n x n block with S symmetric of density `dens`, g ~ N(0,1), rho = 1,
    delta = 1, x free, c_l = c_u on the first m/2 rows, c_l = -inf on the rest, c_k ~ N(0,1).  Returns the shared CSC
    patterns/values and per-instance vectors [batch][len] in the l1-penalty layout of src/QPhandler.cpp:39-51, 150-156."""
    import scipy.sparse as sp
    m = n // 2
    rng = np.random.default_rng(4000 + n if seed is None else seed)
    mask = rng.random((m, n)) < dens
    for i in range(m):
        if not mask[i].any():
            mask[i, rng.integers(0, n)] = True
    J = np.where(mask, rng.standard_normal((m, n)), 0.0)
    U = np.triu(np.where(rng.random((n, n)) < dens, rng.standard_normal((n, n)), 0.0), 1)
    S = U + U.T + np.diag(np.where(rng.random(n) < dens, rng.standard_normal(n), 0.0))
    Hn = S + (np.abs(S).sum(axis=0).max() + 1.0) * np.eye(n)
    nV, nC = n + 2 * m, m
    A = sp.hstack([sp.csc_matrix(J), sp.identity(m, format="csc"), -sp.identity(m, format="csc")], format="csc")
    Hs = sp.block_diag([sp.csc_matrix(Hn), sp.csc_matrix((2 * m, 2 * m))], format="csc")
    A.sort_indices(); Hs.sort_indices()
    Ac = (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64))
    Hc = (Hs.indptr.astype(np.int32), Hs.indices.astype(np.int32), Hs.data.astype(np.float64))
    g = np.concatenate([np.zeros(n), np.ones(2 * m)])[None, :].repeat(batch, 0)
    g[:, :n] = rng.standard_normal((batch, n))
    lb = np.concatenate([np.full(n, -1.0), np.zeros(2 * m)])[None, :].repeat(batch, 0)
    ub = np.concatenate([np.full(n, 1.0), np.full(2 * m, 1e18)])[None, :].repeat(batch, 0)
    ck = rng.standard_normal((batch, m))
    lbA, ubA = -ck.copy(), -ck.copy()
    lbA[:, m // 2:] = -1e18
    return dict(n=n, m=m, nV=nV, nC=nC, Ac=Ac, Hc=Hc, g=np.ascontiguousarray(g), lb=np.ascontiguousarray(lb),
                ub=np.ascontiguousarray(ub), lbA=np.ascontiguousarray(lbA), ubA=np.ascontiguousarray(ubA))
