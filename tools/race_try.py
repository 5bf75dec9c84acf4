import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import restartsqp_b200 as r
from restartsqp_b200 import capi
import helpers as H
name = sys.argv[1] if len(sys.argv) > 1 else 'QORE_hs116'
team = int(sys.argv[2]) if len(sys.argv) > 2 else 0
q = [q for q in H.load_qp_fixtures() if q['name']==name][0]
nV, nC = q['nV'], q['nC']
s = r.CudaQPInterface(nV=nV, nC=nC, qptype=r.QPType.QP, batch=2, team_size=team, keep_state=False)
s.set_csc(capi.MAT_A, q['A_colptr'], q['A_rowidx'], np.array(q['A_val']))
s.set_csc(capi.MAT_H, q['H_colptr'], q['H_rowidx'], np.array(q['H_val']))
s.set_g(np.array(q['g'])); s.set_lb(np.array(q['lb'])); s.set_ub(np.array(q['ub']))
if nC: s.set_lbA(np.array(q['lbA'])); s.set_ubA(np.array(q['ubA']))
s._solve(r.QPType.QP, None, None, 0); s.synchronize()
print(name, s.solve_config(), s.get_status(), s.get_iterations(), s.get_obj_value())
