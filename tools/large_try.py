"""Try the one-QP-per-CTA kernel: small random QPs forced onto it (bit-exact against the oracle), then config-4 shapes."""
import sys, os, time, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
import restartsqp_b200 as r
from restartsqp_b200 import capi
from oracle import oracle_py as orc
import helpers as H

def run(nV, nC, Ac, Hc, g, lb, ub, lbA, ubA, team, label, check=True, maxiter=1000):
    B = g.shape[0]
    s = r.CudaQPInterface(nV=nV, nC=nC, qptype=r.QPType.QP, batch=B, team_size=team, options=r.Options(qp_maxiter=maxiter))
    s.set_csc(capi.MAT_A, *Ac); s.set_csc(capi.MAT_H, *Hc)
    s.set_g(g); s.set_lb(lb); s.set_ub(ub)
    if nC: s.set_lbA(lbA); s.set_ubA(ubA)
    s._solve(r.QPType.QP, None, None, 0); ms = s.last_solve_ms()
    st, it = s.get_status(), s.get_iterations()
    pr = s.profile()
    if pr["total"]:
        print("  profile (% of total cycles): " + "  ".join("%s=%.1f" % (k, 100.0 * v / pr["total"]) for k, v in pr.items() if v and k != "total"), flush=True)
    x = s.get_optimal_solution(); wc, wb = s.get_working_set(translated=False)
    print(f"{label}: nV={nV} nC={nC} B={B} cfg={s.solve_config()} {ms:.2f} ms status={dict(zip(*np.unique(st, return_counts=True)))} iters mean={it.mean():.1f} max={it.max()}", flush=True)
    bad = 0
    if check:
        t0 = time.time()
        for b in range(min(B, check if isinstance(check, int) else B)):
            p = dict(nV=nV, nC=nC, g=g[b], lb=lb[b], ub=ub[b], lbA=lbA[b], ubA=ubA[b])
            o = H.oracle_solve(orc, p, Acsc=Ac, Hcsc=Hc, max_iter=maxiter)
            same = (st[b] == o["status"] and it[b] == o["iters"] and (wb[b] == o["wb"]).all() and (wc[b] == o["wc"]).all())
            err = np.abs(x[b] - o["x"]).max()
            if not same or err > 1e-8: bad += 1; print("  MISMATCH b=%d gpu(st=%d it=%d) oracle(st=%d it=%d) err=%g" % (b, st[b], it[b], o["status"], o["iters"], err))
        print(f"  oracle check: {bad} mismatches, max|dx| last={err:.3g}, oracle {time.time()-t0:.1f}s", flush=True)
    s.close()
    return bad

which = sys.argv[1] if len(sys.argv) > 1 else "small"
rng = np.random.default_rng(5)
if which == "small":
    for (n, m) in [(4, 2), (6, 4), (12, 7), (20, 12), (40, 20)]:
        base = H.random_l1_qp(rng, n, m, convex=True, dens=0.7)
        B = 32
        Ac, Hc = H.csc(base["A"]), H.csc(base["H"])
        g = np.tile(base["g"], (B, 1)); g[:, :n] += rng.standard_normal((B, n))
        t = lambda v: np.ascontiguousarray(np.tile(v, (B, 1)))
        for rep in range(3):
            run(base["nV"], base["nC"], Ac, Hc, g, t(base["lb"]), t(base["ub"]), t(base["lbA"]), t(base["ubA"]), 1024, f"small n={n} rep={rep}")
else:
    n = int(which); B = int(sys.argv[2]) if len(sys.argv) > 2 else 8; chk = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    d = H.synthetic_large_qp(n, batch=B)
    run(d["nV"], d["nC"], d["Ac"], d["Hc"], d["g"], d["lb"], d["ub"], d["lbA"], d["ubA"], 0, f"config4 n={n}", check=chk, maxiter=max(1000, 6 * n))
