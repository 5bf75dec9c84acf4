"""Parity sweep over the whole Hock-Schittkowski suite: the batched SQP loop on the CUDA QP/LP backend against the same loop on
the CPU oracle twin (same host NLP evaluator, same starts): exit flags, outer / QP iteration counts and final iterates."""
import glob, os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from oracle_backend import OracleQPInterface
from test_hs_suite import perturbed_starts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
files = sorted(glob.glob(os.path.join(R, "tests", "golden", "hs_nl", "hs*.nl")))
same = diff = skipped = 0
for k, f in enumerate(files):
    name = os.path.basename(f)[:-3]
    try:
        nlp = AmplNLP(f)
    except NotImplementedError:
        skipped += 1; continue
    if len(nlp.model.G.nodes) > 30000:  # host evaluation of the 1e5-node DAGs takes minutes per iteration
        skipped += 1; print(f"{name:10s} skipped (DAG of {len(nlp.model.G.nodes)} nodes)"); continue
    X = perturbed_starts(nlp, B, k)
    og, oo = r.Options(iter_max=150), r.Options(iter_max=150)
    t0 = time.time()
    rg = BatchedSQP(nlp, x0=X, options=og).Optimize()
    mk = lambda info, qt: r.QPhandler(info, qt, oo, batch=B, backend=OracleQPInterface(info, qt, oo, batch=B), refresh_ubA=True)
    ro = BatchedSQP(nlp, x0=X, options=oo, make_handler=mk).Optimize()
    fin = np.isfinite(ro.x).all(axis=1) & np.isfinite(rg.x).all(axis=1)
    ok = (rg.exitflag == ro.exitflag).all() and (rg.iters == ro.iters).all() and (rg.qp_iter == ro.qp_iter).all() and \
        (np.abs(rg.x[fin] - ro.x[fin]).max() <= 1e-8 * max(1.0, np.abs(ro.x[fin]).max()) if fin.any() else True)
    same += ok; diff += (not ok)
    print(f"{name:10s} n={nlp.n:3d} m={nlp.m:3d} {'identical' if ok else 'DIFFERENT'} flags={dict(zip(*[a.tolist() for a in np.unique(rg.exitflag, return_counts=True)]))} "
          f"iters={int(rg.iters.sum())} qp_iters={int(rg.qp_iter.sum())} {time.time()-t0:.1f}s", flush=True)
    if not ok:
        print("   gpu   ", rg.exitflag, rg.iters, rg.qp_iter); print("   oracle", ro.exitflag, ro.iters, ro.qp_iter)
print(f"TOTAL identical {same}, different {diff}, skipped {skipped}")
