"""The C++ batched host driver (restartsqp_b200/csrc/driver: BatchedAlgorithm, the batched counterpart of the reference's Algorithm
class, and its front end batched_sqp, the counterpart of test/simple_test.cpp).

CPU: the binary is built in-tree, refuses to run without a GPU, and the model hand-over file is exact.  GPU: on the same model
and the same starting points it must reproduce the Python device driver (restartsqp_b200/sqp_device.py) bit for bit -- both only
sequence the same C-ABI calls."""
import os
import subprocess

import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200.nl_reader import AmplNLP, write_model_file
from test_hs_suite import HS_DIR, perturbed_starts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "restartsqp_b200", "lib", "batched_sqp")


def test_driver_binary_is_built_and_has_no_cpu_path(tmp_path):
    assert os.path.exists(BIN), "run python -m restartsqp_b200.build"
    if r.capi.lib().sqpb200_device_count() > 0:
        pytest.skip("a GPU is present")
    host = AmplNLP(os.path.join(HS_DIR, "hs071.nl"))
    path = str(tmp_path / "hs071.model")
    write_model_file(host, path, perturbed_starts(host, 4, 1))
    p = subprocess.run([BIN, path], capture_output=True, text=True)
    assert p.returncode == 2 and "no CUDA device" in p.stderr


def test_model_file_is_exact(tmp_path):
    host = AmplNLP(os.path.join(HS_DIR, "hs071.nl"))
    X = perturbed_starts(host, 5, 2)
    path = str(tmp_path / "m.model")
    write_model_file(host, path, X)
    head, src = open(path).read().split("\n---SOURCE---\n")
    tok = head.split()
    assert [int(t) for t in tok[:5]] == [4, 2, 8, 10, 5]
    vals = [float.fromhex(t) for t in tok[-20:]]
    assert vals == X.ravel().tolist()
    assert "nlp_eval_fc" in src and "nlp_eval_all" in src and src == host.cuda_source()


def run_driver(path, *flags):
    p = subprocess.run([BIN, path, *flags], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = p.stdout.strip().splitlines()
    rows = [ln.split() for ln in lines[:-1]]
    summary = dict(zip(lines[-1].split()[1::2], lines[-1].split()[2::2]))
    ex = np.array([int(t[0]) for t in rows])
    it = np.array([int(t[1]) for t in rows])
    qi = np.array([int(t[2]) for t in rows])
    obj = np.array([float.fromhex(t[3]) for t in rows])
    x = np.array([[float.fromhex(v) for v in t[4:]] for t in rows])
    return ex, it, qi, obj, x, summary


@pytest.mark.gpu
@pytest.mark.parametrize("name,soc", [("hs071", False), ("hs043", True), ("hs038", False)])
def test_cpp_driver_equals_python_device_driver(gpu_lib, tmp_path, name, soc):
    from restartsqp_b200.nl_reader import DeviceNLP
    from restartsqp_b200.sqp_device import DeviceBatchedSQP
    host = AmplNLP(os.path.join(HS_DIR, name + ".nl"))
    B = 96
    X = perturbed_starts(host, B, 4)
    path = str(tmp_path / (name + ".model"))
    write_model_file(host, path, X)
    ex, it, qi, obj, x, summary = run_driver(path, "--iter-max", "120", *(["--soc"] if soc else []))
    dev = DeviceNLP(host)
    alg = DeviceBatchedSQP(dev, x0=X, options=r.Options(iter_max=120, second_order_correction=soc))
    res = alg.Optimize()
    assert (ex == res.exitflag).all() and (it == res.iters).all() and (qi == res.qp_iter).all()
    assert (obj == res.obj).all() and (x == res.x).all()
    assert int(summary["instances"]) == B and int(summary["optimal"]) == int((res.exitflag == 0).sum())
    if name == "hs071":
        assert (ex == 0).mean() > 0.9
        # a second batch on the same object (BatchedAlgorithm::reset): identical results
        ex2, it2, qi2, obj2, x2, _ = run_driver(path, "--iter-max", "120", "--repeat", "2")
        assert (ex2 == ex).all() and (it2 == it).all() and (x2 == x).all()
    alg.close()
    dev.close()
