import sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
from test_hs_suite import HS_DIR, perturbed_starts
name = sys.argv[1] if len(sys.argv) > 1 else "hs100"
dev = DeviceNLP(AmplNLP(os.path.join(HS_DIR, name + ".nl")))
X1, X2 = perturbed_starts(dev.host, 96, 11), perturbed_starts(dev.host, 96, 12)
for pim in (True, False):
    for hs in (False, True):
        a = DeviceBatchedSQP(dev, x0=X1, options=r.Options(iter_max=120), per_instance_modes=pim)
        r1 = a.Optimize(host_sequenced=hs)
        a.reset(X1)
        r1b = a.Optimize(host_sequenced=hs)
        a.reset(X2)
        r2 = a.Optimize(host_sequenced=hs)
        b = DeviceBatchedSQP(dev, x0=X2, options=r.Options(iter_max=120), per_instance_modes=pim)
        r2f = b.Optimize(host_sequenced=hs)
        print("pim", pim, "host_seq", hs, "same-X reset: qp_iter diff", int((r1.qp_iter != r1b.qp_iter).sum()), "x equal", np.array_equal(r1.x, r1b.x),
              "| new-X reset vs fresh: qp_iter diff", int((r2.qp_iter != r2f.qp_iter).sum()), "iters diff", int((r2.iters != r2f.iters).sum()), "x equal", np.array_equal(r2.x, r2f.x))
        a.close(); b.close()
