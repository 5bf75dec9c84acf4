import sys, os, time, cProfile, pstats, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from test_hs_suite import HS_DIR, perturbed_starts
host = AmplNLP(os.path.join(HS_DIR, "hs071.nl")); dev = DeviceNLP(host)
X = perturbed_starts(host, 10000, 0)
warm = BatchedSQP(dev, x0=X[:256]); warm.Optimize()
for rep in range(3):
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter(); alg = BatchedSQP(dev, x0=X); t1 = time.perf_counter(); res = alg.Optimize(); t2 = time.perf_counter()
    pr.disable()
    print("rep", rep, "init %.1f ms  optimize %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
    if rep == 0: pstats.Stats(pr).sort_stats("tottime").print_stats(12)
    alg.myQP_.solverInterface_.close(); alg.myLP_.solverInterface_.close()
