import sys, numpy as np, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import restartsqp_b200 as r
from restartsqp_b200 import capi
from oracle import oracle_py as orc
import helpers as H
print("devices", capi.lib().sqpb200_device_count())
fx = H.load_qp_fixtures()
nbad = 0
for q in fx:
    nV, nC = q['nV'], q['nC']
    B = 3
    s = r.CudaQPInterface(nV=nV, nC=nC, qptype=r.QPType.QP, batch=B)
    s.set_csc(capi.MAT_A, q['A_colptr'], q['A_rowidx'], np.array(q['A_val']))
    s.set_csc(capi.MAT_H, q['H_colptr'], q['H_rowidx'], np.array(q['H_val']))
    s.set_g(np.array(q['g'])); s.set_lb(np.array(q['lb'])); s.set_ub(np.array(q['ub']))
    if nC: s.set_lbA(np.array(q['lbA'])); s.set_ubA(np.array(q['ubA']))
    t0=time.time(); s.optimizeQP(); s.synchronize(); dt=time.time()-t0
    x = s.get_optimal_solution(); st = s.get_status(); it = s.get_iterations(); obj = s.get_obj_value()
    wc, wb = s.get_working_set(translated=False)
    kkt = s.get_optimality_status()['KKT_error']
    A=(q['A_colptr'],q['A_rowidx'],q['A_val']); Hc=(q['H_colptr'],q['H_rowidx'],q['H_val'])
    o = orc.OracleQP(nV,nC); ost = o.init(Hc,q['g'],A,q['lb'],q['ub'],q['lbA'],q['ubA'])
    ox,oy,oobj,oit = o.solution(); owb,owc = o.working_set()
    err = np.abs(x[0]-ox).max()/max(1,np.abs(ox).max())
    same_ws = (wb[0]==owb).all() and (wc[0]==owc).all()
    flag = '' if (st[0]==ost and same_ws and (err<1e-8 or ost!=20)) else '  <<<< MISMATCH'
    if flag: nbad += 1
    print(f"{q['name']:22s} gpu st={st[0]} it={it[0]:4d} obj={obj[0]: .6e} kkt={kkt[0]:.1e} | orc st={ost} it={oit:4d} obj={oobj: .6e} | relerr={err:.1e} ws={same_ws} cfg={s.solve_config()['team_size']} {dt*1e3:.1f}ms{flag}")
    s.close()
print("mismatches", nbad)
