import sys, os, time, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import restartsqp_b200 as r
from restartsqp_b200 import capi
sys.path.insert(0, R)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
team = int(sys.argv[2]) if len(sys.argv) > 2 else 0
only = sys.argv[3] if len(sys.argv) > 3 else None
for k, q in enumerate(bench.load_fixtures()):
    if only and only not in q['name']: continue
    d = bench.make_batch(q, B, 1234 + k)
    t0 = time.time()
    s = r.CudaQPInterface(nV=d["nV"], nC=d["nC"], qptype=r.QPType.QP, batch=B, keep_state=False, team_size=team)
    s.set_csc(capi.MAT_A, q["A_colptr"], q["A_rowidx"], d["Av"]); s.set_csc(capi.MAT_H, q["H_colptr"], q["H_rowidx"], d["Hv"])
    s.set_g(d["g"]); s.set_lb(d["lb"]); s.set_ub(d["ub"])
    if d["nC"]: s.set_lbA(d["lbA"]); s.set_ubA(d["ubA"])
    s.synchronize(); t1 = time.time()
    print(f"{q['name']:20s} nV={d['nV']:3d} nC={d['nC']:3d} cfg={s.solve_config()} setup={t1-t0:.2f}s", end=' ', flush=True)
    for rep in range(2):
        s._solve(r.QPType.QP, None, None, 0); ms = s.last_solve_ms()
    st = s.get_status(); it = s.get_iterations()
    pr = s.profile()
    if pr["total"]:
        print("\n    profile (% of total cycles): " + "  ".join("%s=%.1f" % (k, 100.0 * v / pr["total"]) for k, v in pr.items() if v and k != "total"), end="\n    ")
    print(f"solve={ms:.2f}ms  {B/ms*1e3:.0f} QP/s  status={dict(zip(*np.unique(st, return_counts=True)))} iters mean={it.mean():.1f} max={it.max()}", flush=True)
    s.close()
