"""Parity sweep: the device-resident loop with per-instance init/hotstart decisions (the default of DeviceBatchedSQP) against the C
oracle of the outer loop (oracle/oracle_sqp.c, one independent solve per instance = the reference's semantics), on every HS
problem whose evaluation uses only + - * and squares (so that the NVRTC and the gcc evaluators agree bitwise)."""
import glob, os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
from oracle import oracle_py as orc
from test_hs_suite import perturbed_starts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
same = diff = 0
for k, f in enumerate(sorted(glob.glob(os.path.join(R, "tests", "golden", "hs_nl", "hs*.nl")))):
    try:
        host = AmplNLP(f)
    except NotImplementedError:
        continue
    G = host.model.G
    if any(t[0] in ("sqrt", "sin", "cos", "log", "exp", "abs", "tan", "atan", "tanh", "sinh", "cosh", "log10", "acos", "asin", "div") or
           (t[0] == "pow" and G.cval(t[2]) != 2.0) for t in G.nodes):
        continue
    dev = DeviceNLP(host)
    X = perturbed_starts(host, B, k)
    alg = DeviceBatchedSQP(dev, x0=X, options=r.Options(iter_max=150)); rd = alg.Optimize(); alg.close(); dev.close()
    rc = orc.SqpOracle(host, r.Options(iter_max=150)).solve_batch(X)
    fin = np.isfinite(rc["x"]).all(axis=1) & np.isfinite(rd.x).all(axis=1)
    ok = (rd.exitflag == rc["exitflag"]).all() and (rd.iters == rc["iters"]).all() and (rd.qp_iter == rc["qp_iter"]).all() and np.array_equal(rd.x[fin], rc["x"][fin])
    same += ok; diff += (not ok)
    print(f"{host.name:10s} {'identical' if ok else 'DIFFERENT'} flags={dict(zip(*[a.tolist() for a in np.unique(rd.exitflag, return_counts=True)]))} iters={int(rd.iters.sum())} qp_iters={int(rd.qp_iter.sum())}", flush=True)
    if not ok:
        bad = np.where((rd.exitflag != rc["exitflag"]) | (rd.iters != rc["iters"]) | (rd.qp_iter != rc["qp_iter"]))[0][:4]
        print("   instances", bad, "device", rd.exitflag[bad], rd.iters[bad], rd.qp_iter[bad], "oracle", rc["exitflag"][bad], rc["iters"][bad], rc["qp_iter"][bad])
print(f"TOTAL identical {same}, different {diff}")
