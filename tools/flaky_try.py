import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import restartsqp_b200 as r
from restartsqp_b200 import capi
import helpers as H
name = sys.argv[1]; B = int(sys.argv[2]); reps = int(sys.argv[3])
q = [q for q in H.load_qp_fixtures() if q['name']==name][0]
nV, nC = q['nV'], q['nC']
for team in ([int(t) for t in sys.argv[4].split(',')] if len(sys.argv) > 4 else (32, 64, 128, 256)):
    bad = 0; ref = None
    for rep in range(reps):
        s = r.CudaQPInterface(nV=nV, nC=nC, qptype=r.QPType.QP, batch=B, team_size=team, keep_state=False)
        s.set_csc(capi.MAT_A, q['A_colptr'], q['A_rowidx'], np.array(q['A_val']))
        s.set_csc(capi.MAT_H, q['H_colptr'], q['H_rowidx'], np.array(q['H_val']))
        s.set_g(np.array(q['g'])); s.set_lb(np.array(q['lb'])); s.set_ub(np.array(q['ub']))
        if nC: s.set_lbA(np.array(q['lbA'])); s.set_ubA(np.array(q['ubA']))
        s._solve(r.QPType.QP, None, None, 0)
        st, it, x = s.get_status(), s.get_iterations(), s.get_optimal_solution()
        if ref is None: ref = (st[0], it[0], x[0].copy())
        nb = int(((st != ref[0]) | (it != ref[1]) | (np.abs(x - ref[2]).max(axis=1) != 0)).sum())
        bad += nb
        s.close()
    print(name, "team", team, "ref", ref[0], ref[1], "bad instances", bad, "of", B*reps, flush=True)
