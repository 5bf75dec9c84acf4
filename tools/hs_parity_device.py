"""Parity sweep: the device-resident outer loop (DeviceBatchedSQP, csrc/sqp_outer.cu) against the numpy mirror (BatchedSQP), both
on the NVRTC evaluator and the CUDA QP backend, over every HS problem the device evaluator accepts."""
import glob, os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
from test_hs_suite import perturbed_starts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
files = sorted(glob.glob(os.path.join(R, "tests", "golden", "hs_nl", "hs*.nl")))
same = diff = skipped = 0
for k, f in enumerate(files):
    name = os.path.basename(f)[:-3]
    try:
        dev = DeviceNLP(AmplNLP(f))
    except (NotImplementedError, ValueError):
        skipped += 1; continue
    X = perturbed_starts(dev.host, B, k)
    rh = BatchedSQP(dev, x0=X, options=r.Options(iter_max=150)).Optimize()
    alg = DeviceBatchedSQP(dev, x0=X, options=r.Options(iter_max=150), per_instance_modes=False); rd = alg.Optimize(); alg.close(); dev.close()
    fin = np.isfinite(rh.x).all(axis=1) & np.isfinite(rd.x).all(axis=1)
    ok = (rd.exitflag == rh.exitflag).all() and (rd.iters == rh.iters).all() and (rd.qp_iter == rh.qp_iter).all() and \
        (rd.rho == rh.rho).all() and np.array_equal(rd.x[fin], rh.x[fin])
    same += ok; diff += (not ok)
    print(f"{name:10s} {'identical' if ok else 'DIFFERENT'} flags={dict(zip(*[a.tolist() for a in np.unique(rd.exitflag, return_counts=True)]))} iters={int(rd.iters.sum())} qp_iters={int(rd.qp_iter.sum())}", flush=True)
    if not ok:
        bad = np.where((rd.exitflag != rh.exitflag) | (rd.iters != rh.iters) | (rd.qp_iter != rh.qp_iter))[0][:4]
        print("   first differing instances", bad, "device", rd.exitflag[bad], rd.iters[bad], rd.qp_iter[bad], "host", rh.exitflag[bad], rh.iters[bad], rh.qp_iter[bad])
print(f"TOTAL identical {same}, different {diff}, skipped {skipped}")
