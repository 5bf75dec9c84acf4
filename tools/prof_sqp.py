import sys, os, time, cProfile, pstats, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
from test_hs_suite import HS_DIR, perturbed_starts
host = AmplNLP(os.path.join(HS_DIR, "hs071.nl")); dev = DeviceNLP(host)
X = perturbed_starts(host, 10000, 0)
w = DeviceBatchedSQP(dev, x0=X[:256]); w.Optimize()
for rep in range(3):
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter(); alg = DeviceBatchedSQP(dev, x0=X); t1 = time.perf_counter(); res = alg.Optimize(); t2 = time.perf_counter()
    pr.disable()
    print("rep", rep, "init %.1f ms  optimize %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
    if rep == 1: pstats.Stats(pr).sort_stats("tottime").print_stats(14)
    alg.close()
