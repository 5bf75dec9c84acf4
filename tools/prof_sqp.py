import sys, os, time, cProfile, pstats, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from test_hs_suite import HS_DIR, perturbed_starts
name = sys.argv[1] if len(sys.argv) > 1 else "hs071"; B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
host = AmplNLP(os.path.join(HS_DIR, name + ".nl")); dev = DeviceNLP(host)
X = perturbed_starts(host, B, 0)
BatchedSQP(dev, x0=X[:256]).Optimize()
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
res = BatchedSQP(dev, x0=X).Optimize()
dt = time.perf_counter() - t0; pr.disable()
print(name, B, "solves/s", B / dt, "optimal", (res.exitflag == 0).sum(), "iters", res.iters.mean(), res.iters.max())
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
