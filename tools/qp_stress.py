"""Randomised parity stress of the QP/LP kernels against the oracle: random l1-penalty QPs (convex and non-convex, random shapes
and densities), cold start, then two hot starts with new vectors and one with new matrix values, on the warp kernel (with a small
factor capacity so that the rescue launch is exercised) and on the one-QP-per-CTA kernel.  The warp kernel is compared bitwise; the CTA kernel (DMMA refactorisation) at north_star's 1e-8 gate."""
import sys, os, time, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200 import capi
from oracle import oracle_py as orc
import helpers as H
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = tot = 0
t0 = time.time()
for case in range(ncase):
    n, m = int(rng.integers(1, 26)), int(rng.integers(0, 16))
    convex = bool(rng.random() < 0.7); is_lp = bool(rng.random() < 0.15)
    base = H.random_l1_qp(rng, n, m, convex=convex, dens=float(rng.uniform(0.2, 1.0)), rho=float(10.0 ** rng.integers(0, 4)))
    nV, nC, B = base["nV"], base["nC"], 6
    Ac, Hc = H.csc(base["A"]), H.csc(base["H"])
    t = lambda v: np.ascontiguousarray(np.tile(v, (B, 1)))
    for team, cap in ((0, 0), (0, max(1, n // 2)), (1024, 0)):
        g = t(base["g"]); g[:, :n] += rng.standard_normal((B, n))
        if is_lp: g[:, :n] = 0.0
        lb, ub, lbA, ubA = t(base["lb"]), t(base["ub"]), t(base["lbA"]), t(base["ubA"])
        qt = r.QPType.LP if is_lp else r.QPType.QP
        s = r.CudaQPInterface(nV=nV, nC=nC, qptype=qt, batch=B, team_size=team, factor_cap=cap)
        s.set_csc(capi.MAT_A, *Ac)
        if not is_lp: s.set_csc(capi.MAT_H, *Hc)
        Av, Hv = t(Ac[2]), t(Hc[2])
        sol = [orc.OracleQP(nV, nC, max_iter=100 if is_lp else 1000) for _ in range(B)]
        for step in range(5):  # init, 2 x hotstart (FIXED), init from the previous solution (FIXED -> VARIED flip), hotstart with matrices
            if step in (1, 2):
                g = g.copy(); g[:, :n] += 0.3 * rng.standard_normal((B, n)) * (0.0 if is_lp else 1.0)
                if m: lbA = np.where(lbA > -1e17, lbA + 0.2 * rng.standard_normal((B, m)), lbA); ubA = np.maximum(ubA, lbA)
            if step >= 3:
                Av = Av * (1.0 + 0.05 * rng.standard_normal(Av.shape) * (np.abs(np.abs(Av) - 1.0) > 1e-12))
                s.set_csc_values(capi.MAT_A, Av)
                if not is_lp:
                    Hv = Hv * 1.05; s.set_csc_values(capi.MAT_H, Hv)
            s.set_g(g); s.set_lb(lb); s.set_ub(ub)
            if m: s.set_lbA(lbA); s.set_ubA(ubA)
            s._solve(qt, None, None, 0)
            x, st, it = s.get_optimal_solution(), s.get_status(), s.get_iterations()
            y = np.concatenate([s.get_multipliers_bounds(), s.get_multipliers_constr()], axis=1)
            for b in range(B):
                o = sol[b]
                if step == 0 or not o_ok[b]:
                    so = o.init(None if is_lp else (Hc[0], Hc[1], Hv[b]), g[b], (Ac[0], Ac[1], Av[b]), lb[b], ub[b], lbA[b], ubA[b], is_lp=is_lp)
                    ito = o.solution()[3]
                else:
                    so = o.hotstart(g[b], lb[b], ub[b], lbA[b], ubA[b]) if step < 3 else (o.reinit if step == 3 else o.hotstart_matrices)(None if is_lp else Hv[b], Av[b], g[b], lb[b], ub[b], lbA[b], ubA[b])
                    ito = o.solution()[3]
                if so != 20:  # handle_error, after an init as well as after a hot start
                    so, added = o.handle_error()
                    ito += added
                xo, yo, _, _ = o.solution()
                tot += 1
                if team != 1024:  # warp kernel: bit for bit
                    same = so == int(st[b]) and ito == int(it[b]) and (so != 20 or (np.array_equal(xo, x[b], equal_nan=True) and np.array_equal(yo, y[b], equal_nan=True)))
                elif is_lp:  # CTA kernel (DMMA sums, block-inverse solves): north_star's gate -- same status, same optimum to 1e-8
                    same = so == int(st[b]) and (so != 20 or abs(g[b] @ xo - g[b] @ x[b]) <= 1e-8 * max(1.0, abs(g[b] @ xo)))
                elif convex:
                    rel = lambda a_, b_: np.abs(a_ - b_).max() / max(1.0, np.abs(b_).max())
                    same = so == int(st[b]) and (so != 20 or (rel(x[b], xo) <= 1e-8 and rel(y[b], yo) <= 1e-7))
                else:  # non-convex: the path (and the local solution reached) is rounding-sensitive; require a clean exit
                    same = 20 <= int(st[b]) <= 30
                if not same:
                    bad += 1
                    if bad <= 10: print(f"MISMATCH case {case} (n={n} m={m} convex={convex} lp={is_lp}) team={team} cap={cap} step={step} b={b}: gpu st={st[b]} it={it[b]}  oracle st={so} it={ito}  max|dx|={np.abs(xo - x[b]).max():.3g}", flush=True)
            if step == 0: o_ok = np.zeros(B, bool)
            o_ok = np.array([int(st[b]) == 20 for b in range(B)])
        s.close()
print(f"{tot} solves compared, {bad} mismatches, {time.time()-t0:.0f}s")
