"""Run the Hock-Schittkowski .nl suite (tests/golden/hs_nl, the reference's test/CUTE_examples/hs*.nl) through the batched SQP.
  python tools/hs_suite.py [--backend cuda|oracle] [--batch B] [--only hs071,hs035] [--iter-max N]
Start points: SURVEY.md 8d config 3 (x0_i = clip(x0*(1+0.1 N) + 0.1 N), seed 71000 + problem index; instance 0 = x0)."""
import argparse, glob, os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP
from restartsqp_b200.sqp_driver import BatchedSQP


def starts(nlp, B, k):
    x0, _ = nlp.Get_starting_point()
    xl, xu, _, _ = nlp.Get_bounds_info()
    rng = np.random.default_rng(71000 + k)
    X = np.clip(x0 * (1 + 0.1 * rng.standard_normal((B, nlp.n))) + 0.1 * rng.standard_normal((B, nlp.n)), xl, xu)
    X[0] = np.clip(x0, xl, xu)
    return X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="cuda"); ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--only", default=""); ap.add_argument("--iter-max", type=int, default=300)
    ap.add_argument("--loop", default="host", choices=["host", "device"], help="device: DeviceNLP + DeviceBatchedSQP (iterates stay on the GPU)")
    a = ap.parse_args()
    files = sorted(glob.glob(os.path.join(R, "tests", "golden", "hs_nl", "hs*.nl")))
    only = set(a.only.split(",")) if a.only else None
    tot = dict(inst=0, opt=0, t=0.0)
    for k, f in enumerate(files):
        name = os.path.splitext(os.path.basename(f))[0]
        if only and name not in only:
            continue
        try:
            nlp = AmplNLP(f)
        except NotImplementedError as e:
            print(f"{name:10s} skipped: {e}"); continue
        opt = r.Options(iter_max=a.iter_max)
        mk = None
        if a.backend == "oracle":
            from oracle_backend import OracleQPInterface
            mk = lambda info, qptype: r.QPhandler(info, qptype, opt, batch=a.batch, backend=OracleQPInterface(info, qptype, opt, batch=a.batch), refresh_ubA=True)
        tc = 0.0
        if a.loop == "device":
            from restartsqp_b200.nl_reader import DeviceNLP
            from restartsqp_b200.sqp_device import DeviceBatchedSQP
            try:
                t0 = time.time(); dnlp = DeviceNLP(nlp); tc = time.time() - t0
            except ValueError as e:
                print(f"{name:10s} n={nlp.n:3d} m={nlp.m:3d} skipped (device evaluator): {str(e)[:90]}"); continue
        X = starts(nlp, a.batch, k)
        t0 = time.time()
        try:
            if a.loop == "device":
                alg = DeviceBatchedSQP(dnlp, x0=X, options=opt)
                res = alg.Optimize()
                alg.close(); dnlp.close()
            else:
                alg = BatchedSQP(nlp, x0=X, options=opt, make_handler=mk)
                res = alg.Optimize()
        except Exception as e:
            print(f"{name:10s} n={nlp.n:3d} m={nlp.m:3d} ERROR {type(e).__name__}: {str(e)[:100]}"); continue
        dt = time.time() - t0
        fl, cnt = np.unique(res.exitflag, return_counts=True)
        nopt = int((res.exitflag == int(r.Exitflag.OPTIMAL)).sum())
        tot["inst"] += a.batch; tot["opt"] += nopt; tot["t"] += dt
        print(f"{name:10s} n={nlp.n:3d} m={nlp.m:3d} optimal {nopt}/{a.batch} flags={dict(zip(fl.tolist(), cnt.tolist()))} "
              f"iters mean={res.iters.mean():.1f} qp_iter mean={res.qp_iter.mean():.1f} f[0]={res.obj[0]:.6g} {dt:.3f}s {a.batch / dt:.0f} solves/s nvrtc {tc:.1f}s", flush=True)
    print(f"TOTAL optimal {tot['opt']}/{tot['inst']} in {tot['t']:.1f}s")


if __name__ == "__main__":
    main()
