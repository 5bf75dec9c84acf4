"""Randomised bitwise check of the staged L0 kernels (TMA SpMV, TMA KKT test) against the oracle's restatement of the reference
formulas: random shapes (odd / even nnz and vector lengths, nC = 0), batches that leave odd tail groups, per-instance values."""
import sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200 import capi
from oracle import oracle_py as orc
import helpers as H
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
bad = tot = 0
for case in range(ncase):
    n, m = int(rng.integers(1, 40)), int(rng.integers(0, 25))
    B = int(rng.choice([1, 2, 3, 5, 8, 31, 64, 257, 1001]))
    base = H.random_l1_qp(rng, n, m, convex=True, dens=float(rng.uniform(0.1, 1.0)))
    nV, nC = base["nV"], base["nC"]
    Ac, Hc = H.csc(base["A"]), H.csc(base["H"])
    Av = np.tile(Ac[2], (B, 1)) * (1 + 0.1 * rng.standard_normal((B, len(Ac[2]))))
    Hv = np.tile(Hc[2], (B, 1)) * (1 + 0.1 * rng.random((B, 1)))  # per-instance scaling: H stays symmetric (the kernels read row c as column c)
    s = r.CudaQPInterface(nV=nV, nC=nC, qptype=r.QPType.QP, batch=B)
    s.set_csc(capi.MAT_A, Ac[0], Ac[1], Av); s.set_csc(capi.MAT_H, Hc[0], Hc[1], Hv)
    x, yc = rng.standard_normal((B, nV)), rng.standard_normal((B, nC))
    Ax = s.spmv(capi.MAT_A, x) if nC else np.zeros((B, 0))
    ATy = s.spmv(capi.MAT_A, yc, transpose=True)
    Hx = s.spmv(capi.MAT_H, x)
    # a solve so that the handle holds x, y and a working set; then the stand-alone KKT kernel on them
    g = np.tile(base["g"], (B, 1)); g[:, :n] += rng.standard_normal((B, n))
    t = lambda v: np.ascontiguousarray(np.tile(v, (B, 1)))
    lb, ub, lbA, ubA = t(base["lb"]), t(base["ub"]), t(base["lbA"]), t(base["ubA"])
    s.set_g(g); s.set_lb(lb); s.set_ub(ub)
    if nC: s.set_lbA(lbA); s.set_ubA(ubA)
    s._solve(r.QPType.QP, None, None, 0)
    k = s.get_optimality_status(recompute=True)
    xs = s.get_optimal_solution(); ys = np.concatenate([s.get_multipliers_bounds(), s.get_multipliers_constr()], axis=1)
    wc, wb = s.get_working_set(translated=False)
    WcT, WbT = s.get_working_set(translated=True)
    for b in set([0, B - 1, B // 2, int(rng.integers(0, B))]):
        A_b, H_b = (Ac[0], Ac[1], Av[b]), (Hc[0], Hc[1], Hv[b])
        ok = True
        if nC: ok &= np.array_equal(Ax[b], orc.csc_times(nC, nV, *A_b, x[b]))
        ok &= np.array_equal(ATy[b], orc.csc_times(nC, nV, *A_b, yc[b], transpose=True)) if nC else not ATy[b].any()
        ok &= np.array_equal(Hx[b], orc.csc_times(nV, nV, *H_b, x[b]))
        Axs = orc.csc_times(nC, nV, *A_b, xs[b]) if nC else np.zeros(0)
        Wb, Wc = orc.translate_working_set(wb[b], wc[b], xs[b], Axs, lb[b], ub[b], lbA[b], ubA[b])
        _, res = orc.kkt_residuals(nV, nC, A_b, H_b, g[b], lb[b], ub[b], lbA[b], ubA[b], xs[b], ys[b], Wb, Wc)
        got = [k[kk][b] for kk in ("primal_violation", "dual_violation", "stationarity_violation", "compl_violation", "KKT_error")]
        ok &= got == res.tolist() and np.array_equal(WbT[b], Wb) and np.array_equal(WcT[b], Wc)
        tot += 1
        if not ok:
            bad += 1
            if bad <= 8: print(f"MISMATCH case {case} n={n} m={m} B={B} b={b} kkt gpu={got} oracle={res.tolist()}", flush=True)
    s.close()
print(f"{tot} instances compared, {bad} mismatches")
