"""HS071 x B perturbed starts through the device-resident SQP loop (one C call): wall-clock split and launch count."""
import sys, os, time, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
name = sys.argv[1] if len(sys.argv) > 1 else "hs071"
Bs = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
host = AmplNLP(os.path.join(R, "tests", "golden", "hs_nl", name + ".nl")); dev = DeviceNLP(host)
x0, _ = host.Get_starting_point(); xl, xu, _, _ = host.Get_bounds_info()
rng = np.random.default_rng(71000)
X = np.clip(x0 * (1 + 0.1 * rng.standard_normal((Bs, host.n))) + 0.1 * rng.standard_normal((Bs, host.n)), xl, xu)
DeviceBatchedSQP(dev, x0=X[:256]).Optimize()
for rep in range(reps):
    t0 = time.perf_counter(); alg = DeviceBatchedSQP(dev, x0=X); t1 = time.perf_counter(); res = alg.Optimize(); t2 = time.perf_counter()
    print(name, Bs, "init %.2f ms  optimize %.2f ms  total %.2f ms  -> %.2f M solves/s  optimal %d launches %d iters mean %.2f" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), Bs / (t2 - t0) / 1e6, (res.exitflag == 0).sum(), alg.launches, res.iters.mean()), flush=True)
    for rr in range(3):  # the same object serving further batches: reset (upload + evaluation + state) and Optimize
        Xr = np.clip(x0 * (1 + 0.1 * rng.standard_normal((Bs, host.n))) + 0.1 * rng.standard_normal((Bs, host.n)), xl, xu)
        t0 = time.perf_counter(); alg.reset(Xr); t1 = time.perf_counter(); res = alg.Optimize(); t2 = time.perf_counter()
        print("   reuse: reset %.2f ms  optimize %.2f ms  total %.2f ms  -> %.2f M solves/s  optimal %d" % (
            1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), Bs / (t2 - t0) / 1e6, (res.exitflag == 0).sum()), flush=True)
    alg.close()
