"""The C++ drop-in plugin (restartsqp_b200/csrc/adapter/CudaQPInterface.{hpp,cpp}).

CPU: it compiles against the reference's own headers (only where /root/reference exists) and the prebuilt driver
refuses to run without a GPU.  GPU: the driver oracle/_ref/adapter_hs071 (the plugin + the reference's own
Vector/SpTripletMat/SpHbMat/Options classes, linked with libsqpb200.so) solves the first HS071 subproblem the
way QPhandler would and must agree with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as orc
import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "adapter_hs071")


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="needs the reference headers")
def test_adapter_compiles_against_reference_headers():
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-w", "-I" + os.path.join(ROOT, "oracle", "stubs"), "-I/root/reference/include",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "restartsqp_b200", "csrc", "adapter", "CudaQPInterface.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="needs the reference headers")
def test_qore_layout_adapter_compiles_against_reference_headers():
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-w", "-I" + os.path.join(ROOT, "oracle", "stubs"), "-I/root/reference/include",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "restartsqp_b200", "csrc", "adapter", "CudaQOREInterface.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_qore_layout_adapter_implements_every_pure_virtual():
    hpp = open(os.path.join(ROOT, "restartsqp_b200", "csrc", "adapter", "CudaQOREInterface.hpp")).read()
    assert hpp.count("override") >= 30  # the 30 pure virtuals of include/sqphot/QPsolverInterface.hpp:43-194
    for ref in ("include/sqphot/QOREInterface.hpp:30-252", "src/QPhandler.cpp:225-260"):
        assert ref in hpp


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="needs the reference tree")
def test_reference_patch_applies(tmp_path):
    """integration/restartsqp_cuda_backend.patch (the registration of the two plugins) applies to the
    reference tree as it is; the patched Options.cpp and both plugins compile against the patched headers."""
    import shutil
    ref = "/root/reference"
    work = str(tmp_path / "ref")
    for f in ("include/sqphot", "src"):
        shutil.copytree(os.path.join(ref, f), os.path.join(work, f))
    for dirpath, _, files in os.walk(work):
        os.chmod(dirpath, 0o755)
        for f in files:
            os.chmod(os.path.join(dirpath, f), 0o644)
    patch = os.path.join(ROOT, "integration", "restartsqp_cuda_backend.patch")
    p = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", patch], cwd=work, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "CUDA_B200_QORE_LAYOUT" in open(os.path.join(work, "include/sqphot/Types.hpp")).read()
    adapter = os.path.join(ROOT, "restartsqp_b200", "csrc", "adapter")
    for f in ("CudaQPInterface", "CudaQOREInterface"):
        shutil.copy(os.path.join(adapter, f + ".hpp"), os.path.join(work, "include/sqphot", f + ".hpp"))
        shutil.copy(os.path.join(adapter, f + ".cpp"), os.path.join(work, "src", f + ".cpp"))
    inc = ["-I" + os.path.join(ROOT, "oracle", "stubs"), "-I" + os.path.join(work, "include"), "-I" + os.path.join(work, "include/sqphot"),
           "-I" + os.path.join(ROOT, "include")]
    for src in ("src/Options.cpp", "src/CudaQPInterface.cpp", "src/CudaQOREInterface.cpp"):
        c = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-w"] + inc + [os.path.join(work, src)], capture_output=True, text=True)
        assert c.returncode == 0, src + "\n" + c.stderr


def test_adapter_implements_every_pure_virtual():
    hpp = open(os.path.join(ROOT, "restartsqp_b200", "csrc", "adapter", "CudaQPInterface.hpp")).read()
    for name in ["getLb", "getUb", "getLbA", "getUbA", "getG", "getH", "getA", "optimizeQP", "optimizeLP", "get_optimal_solution",
                 "get_obj_value", "get_multipliers_bounds", "get_multipliers_constr", "get_working_set", "get_status",
                 "test_optimality", "get_optimality_status", "set_lb", "set_ub", "set_lbA", "set_ubA", "set_g", "set_H", "set_A",
                 "reset_constraints", "WriteQPDataToFile"]:
        assert name in hpp, name
    assert hpp.count("override") >= 30  # the 30 pure virtuals of include/sqphot/QPsolverInterface.hpp:43-194


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/adapter_hs071 not built")
def test_driver_refuses_to_run_without_gpu():
    import restartsqp_b200 as r
    if r.capi.lib().sqpb200_device_count() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([DRIVER], capture_output=True, text=True)
    assert p.returncode == 2 and "create_failed" in p.stdout


def hs071_first_qp(delta):
    n, m = 4, 2
    xk = np.array([1.0, 5.0, 5.0, 1.0])
    jr, jc = [1, 2] * 4, [1, 1, 2, 2, 3, 3, 4, 4]
    J = np.array([[xk[1] * xk[2] * xk[3], xk[0] * xk[2] * xk[3], xk[0] * xk[1] * xk[3], xk[0] * xk[1] * xk[2]], 2 * xk])
    jv = np.array([J[a - 1, b - 1] for a, b in zip(jr, jc)])
    hr, hc = [1, 1, 2, 1, 2, 3, 1, 2, 3, 4], [1, 2, 2, 3, 3, 3, 4, 4, 4, 4]
    Hd = np.array([[2 * xk[3] + 2, xk[3], xk[3], 2 * xk[0] + xk[1] + xk[2]], [0, 2, 0, xk[0]], [0, 0, 2, xk[0]], [0, 0, 0, 2.0]])
    hv = np.array([Hd[a - 1, b - 1] for a, b in zip(hr, hc)])
    A = orc.assemble_A(m, n + 2 * m, jr, jc, jv, orc.identity_info(n, m))
    Hh = orc.assemble_H(n + 2 * m, hr, hc, hv, True)
    ck = np.array([xk.prod(), (xk ** 2).sum()])
    lb, ub, lbA, ubA = np.zeros(8), np.zeros(8), np.zeros(2), np.zeros(2)
    orc.qp_bounds(0, n, m, delta, np.ones(4), 5 * np.ones(4), xk, np.array([25.0, 40.0]), np.array([1e18, 40.0]), ck, lb, ub, lbA, ubA)
    grad = np.array([xk[3] * (2 * xk[0] + xk[1] + xk[2]), xk[0] * xk[3], xk[0] * xk[3] + 1, xk[0] * (xk[0] + xk[1] + xk[2])])
    g = np.concatenate([grad, np.ones(4)])
    return dict(nV=8, nC=2, g=g, lb=lb, ub=ub, lbA=lbA, ubA=ubA), A, Hh


@pytest.mark.gpu
def test_cpp_plugin_matches_oracle_on_hs071(gpu_lib):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/adapter_hs071 not built (needs /root/reference at build time)")
    p = subprocess.run([DRIVER], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    out = {}
    for line in p.stdout.strip().splitlines():
        k, *v = line.split()
        out[k] = v
    prob, A, Hh = hs071_first_qp(1.0)
    o = H.oracle_solve(orc, prob, Acsc=A[:3], Hcsc=Hh[:3])
    assert int(out["status"][0]) == 20 == o["status"] and int(out["kkt_ok"][0]) == 1
    assert int(out["qp_iter"][0]) == o["iters"]
    assert np.array(out["x"], float).tolist() == o["x"].tolist()  # bit-identical
    assert np.array(out["y"], float).tolist() == o["y"].tolist()
    assert float(out["obj"][0]) == o["obj"]
    assert [int(t) for t in out["A_colptr"]] == A[0].tolist() and [int(t) for t in out["A_rowidx"]] == A[1].tolist()
    Ax = orc.csc_times(2, 8, A[0], A[1], A[2], o["x"])
    Wb, Wc = orc.translate_working_set(o["wb"], o["wc"], o["x"], Ax, prob["lb"], prob["ub"], prob["lbA"], prob["ubA"])
    assert [int(t) for t in out["Wb"]] == Wb.tolist() and [int(t) for t in out["Wc"]] == Wc.tolist()
    # hot start after update_delta(0.5): same point as the oracle's hotstart
    prob2, _, _ = hs071_first_qp(0.5)
    st = o["solver"].hotstart(prob2["g"], prob2["lb"], prob2["ub"], prob2["lbA"], prob2["ubA"])
    x2, y2, obj2, it2 = o["solver"].solution()
    assert int(out["hot_status"][0]) == st == 20
    assert np.array(out["hot_x"], float).tolist() == x2.tolist()
    assert int(out["hot_qp_iter"][0]) == o["iters"] + it2


@pytest.mark.gpu
def test_cpp_qore_layout_plugin_matches_oracle_on_hs071(gpu_lib):
    """The QORE-layout plugin (CudaQOREInterface.cpp) driven like QPhandler's QORE branch: x_qp = [x ; A x], one stacked
    multiplier vector, QORE's working-set sign, row-compressed getA(); same numbers as the oracle, bit for bit."""
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/adapter_hs071 not built (needs /root/reference at build time)")
    p = subprocess.run([DRIVER, "qore"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    out = {}
    for line in p.stdout.strip().splitlines():
        k, *v = line.split()
        out[k] = v
    prob, A, Hh = hs071_first_qp(1.0)
    o = H.oracle_solve(orc, prob, Acsc=A[:3], Hcsc=Hh[:3])
    assert int(out["status"][0]) == 20 == o["status"] and int(out["kkt_ok"][0]) == 1 and int(out["getLbA_threw"][0]) == 1
    assert int(out["qp_iter"][0]) == o["iters"]
    Ax = orc.csc_times(2, 8, A[0], A[1], A[2], o["x"])
    assert np.array(out["x"], float).tolist() == o["x"].tolist() + Ax.tolist()  # [x ; A x], bit-identical
    assert np.array(out["y"], float).tolist() == o["y"].tolist()
    assert float(out["obj"][0]) == o["obj"]
    assert [int(t) for t in out["ws"]] == (-np.concatenate([o["wb"], o["wc"]])).tolist()  # QORE: -1 upper, +1 lower
    Wb, Wc = orc.translate_working_set(o["wb"], o["wc"], o["x"], Ax, prob["lb"], prob["ub"], prob["lbA"], prob["ubA"])
    assert [int(t) for t in out["Wb"]] == Wb.tolist() and [int(t) for t in out["Wc"]] == Wc.tolist()
    # getA() is the reference's compressed-row matrix of the same triplets
    jr, jc = [1, 2] * 4, [1, 1, 2, 2, 3, 3, 4, 4]
    xk = np.array([1.0, 5.0, 5.0, 1.0])
    J = np.array([[xk[1] * xk[2] * xk[3], xk[0] * xk[2] * xk[3], xk[0] * xk[1] * xk[3], xk[0] * xk[1] * xk[2]], 2 * xk])
    rp, ci, v, _ = orc.assemble_A_csr(2, 8, jr, jc, np.array([J[a - 1, b - 1] for a, b in zip(jr, jc)]), orc.identity_info(4, 2))
    assert [int(t) for t in out["A_rowptr"]] == rp.tolist() and [int(t) for t in out["A_colidx"]] == ci.tolist()
    assert np.array(out["A_val"], float).tolist() == v.tolist()
    assert np.array(out["lb"], float).tolist() == prob["lb"].tolist() + prob["lbA"].tolist()
    prob2, _, _ = hs071_first_qp(0.5)
    st = o["solver"].hotstart(prob2["g"], prob2["lb"], prob2["ub"], prob2["lbA"], prob2["ubA"])
    x2, y2, obj2, it2 = o["solver"].solution()
    assert int(out["hot_status"][0]) == st == 20
    assert np.array(out["hot_x"], float)[:8].tolist() == x2.tolist()
    assert int(out["hot_qp_iter"][0]) == o["iters"] + it2
