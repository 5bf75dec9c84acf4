import sys, os, time, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
from test_hs_suite import HS_DIR, perturbed_starts
name = sys.argv[1] if len(sys.argv) > 1 else "hs071"; B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
host = AmplNLP(os.path.join(HS_DIR, name + ".nl")); dev = DeviceNLP(host)
X = perturbed_starts(host, B, 0)
w = DeviceBatchedSQP(dev, x0=X[:256]); w.Optimize(); w.close()
for rep in range(2):
    t0 = time.perf_counter(); alg = DeviceBatchedSQP(dev, x0=X); t1 = time.perf_counter(); res = alg.Optimize(); t2 = time.perf_counter()
    print(f"device loop {name} B={B}: init {1e3*(t1-t0):.1f} ms, optimize {1e3*(t2-t1):.1f} ms, {B/(t2-t0):.0f} solves/s, optimal {(res.exitflag==0).sum()}, launches {alg.launches}", flush=True)
    alg.close()
if B <= 200000:
    t0 = time.perf_counter(); res_h = BatchedSQP(dev, x0=X).Optimize(); t2 = time.perf_counter()
    print(f"host loop: {1e3*(t2-t0):.1f} ms; same exitflags {(res_h.exitflag==res.exitflag).all()} same x {np.array_equal(res_h.x, res.x)} iters {(res_h.iters==res.iters).all()}")
