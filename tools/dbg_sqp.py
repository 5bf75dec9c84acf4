"""Debug aid: run one HS problem through the SQP loop on the CUDA backend and on the oracle twin, log every QP/LP solve
(per-instance iterations, status, x) and print the first call where they differ."""
import sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from restartsqp_b200 import qp_handler
from oracle_backend import OracleQPInterface
from test_hs_suite import HS_DIR, perturbed_starts

name = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 6; k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nlp = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
X = perturbed_starts(nlp, B, k)
log = {}

def run(tag, mk):
    L = log[tag] = []
    oq, ol = qp_handler.QPhandler.solveQP, qp_handler.QPhandler.solveLP
    def sq(self, stats=None, options=None, active_mask=None):
        out = oq(self, stats, options, active_mask)
        si = self.solverInterface_
        L.append(("QP", None if active_mask is None else np.array(active_mask).copy(), si.get_iterations().copy(), si.get_status().copy(), si.get_optimal_solution().copy(), si.getG().copy(), si.getLb().copy(), si.getUb().copy(), si.getLbA().copy(), si.getUbA().copy()))
        return out
    def sl(self, stats=None, active_mask=None):
        out = ol(self, stats, active_mask)
        si = self.solverInterface_
        L.append(("LP", None if active_mask is None else np.array(active_mask).copy(), si.get_iterations().copy(), si.get_status().copy(), si.get_optimal_solution().copy(), si.getG().copy(), si.getLb().copy(), si.getUb().copy(), si.getLbA().copy(), si.getUbA().copy()))
        return out
    qp_handler.QPhandler.solveQP, qp_handler.QPhandler.solveLP = sq, sl
    try:
        opt = r.Options(iter_max=150)
        res = BatchedSQP(nlp, x0=X, options=opt, make_handler=mk(opt)).Optimize()
    finally:
        qp_handler.QPhandler.solveQP, qp_handler.QPhandler.solveLP = oq, ol
    return res

rg = run("gpu", lambda opt: None)
ro = run("orc", lambda opt: (lambda info, qt: r.QPhandler(info, qt, opt, batch=B, backend=OracleQPInterface(info, qt, opt, batch=B), refresh_ubA=True)))
print("exit", rg.exitflag, ro.exitflag); print("qp_iter", rg.qp_iter, ro.qp_iter)
for c, (a, b) in enumerate(zip(log["gpu"], log["orc"])):
    m = np.ones(B, bool) if a[1] is None else a[1].astype(bool)
    same_in = all(np.array_equal(a[j][m], b[j][m]) for j in range(5, 10))
    if a[0] != b[0] or not np.array_equal(a[2][m], b[2][m]) or not np.array_equal(a[3][m], b[3][m]) or not np.array_equal(a[4][m], b[4][m]) or not same_in:
        print("first difference at call", c, a[0], b[0], "mask", m.astype(int), "inputs identical:", same_in)
        print(" gpu iters", a[2], "status", a[3]); print(" orc iters", b[2], "status", b[3])
        print(" max |x_gpu - x_orc| per instance", np.abs(a[4] - b[4]).max(axis=1))
        for j, nm in zip(range(5, 10), ("g", "lb", "ub", "lbA", "ubA")):
            if a[j].size: print("  input", nm, "max diff per instance", np.abs(a[j] - b[j]).max(axis=1))
        bad = np.where(m & ((a[2] != b[2]) | (a[3] != b[3])))[0]
        for i in bad[:2]:
            print(" inst", i, "x gpu", a[4][i], "\n        x orc", b[4][i])
        break
else:
    print("no difference in", len(log["gpu"]), "calls")
print("calls", len(log["gpu"]), len(log["orc"]))
