"""GPU parity for the L0 rows (A4-A7), the QPhandler data kernels (B2, B3) and the stand-alone KKT kernel
(C7, C8).  Everything here is bit-exact: integer arrays and FP64 results are compared with ==, against the
committed golden vectors produced by the reference's own code and against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200 import capi
from oracle import oracle_py as orc
import helpers as H

pytestmark = pytest.mark.gpu
CASES = H.load_l0_golden()


def make_iface(case, batch):
    n, m = case["n"], case["m"]
    info = r.NLPInfo(nCon=m, nVar=n)
    s = r.CudaQPInterface(info, r.QPType.QP, batch=batch)
    I = orc.identity_info(n, m)
    J = r.SpTripletMat(np.array(case["J_row1"], np.int32), np.array(case["J_col1"], np.int32), np.array(case["J_val"]), m, n)
    Hm = r.SpTripletMat(np.array(case["H_row1"], np.int32), np.array(case["H_col1"], np.int32), np.array(case["H_val"]),
                        n + 2 * m, n + 2 * m, True)
    s.set_A(J, r.IdentityInfo(*I))
    s.set_H(Hm)
    return s, J, Hm


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_assembly_bit_exact_with_reference(gpu_lib, case):
    s, J, Hm = make_iface(case, batch=3)
    A, Hh = s.getA(), s.getH()
    assert A["ColIndex"].tolist() == case["A_colptr"]
    assert A["RowIndex"].tolist() == case["A_rowidx"]
    assert A["order"].tolist() == case["A_order"]
    assert Hh["ColIndex"].tolist() == case["H_colptr"]
    assert Hh["RowIndex"].tolist() == case["H_rowidx"]
    assert Hh["order"].tolist() == case["H_order"]
    for b in range(3):
        assert A["MatVal"][b].tolist() == case["A_val"]
        assert Hh["MatVal"][b].tolist() == case["H_cscval"]
    # value refresh through `order` (SpHbMat::setMatVal), per-instance values
    J.MatVal = np.stack([np.array(case["J_val2"]), np.array(case["J_val"]), 2.0 * np.array(case["J_val2"])])
    Hm.MatVal = np.stack([np.array(case["H_val2"]), np.array(case["H_val"]), 2.0 * np.array(case["H_val2"])])
    s.set_A(J, None)
    s.set_H(Hm)
    A, Hh = s.getA(), s.getH()
    assert A["MatVal"][0].tolist() == case["A_val2"] and A["MatVal"][1].tolist() == case["A_val"]
    assert Hh["MatVal"][0].tolist() == case["H_cscval2"] and Hh["MatVal"][1].tolist() == case["H_cscval"]
    s.close()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_spmv_bit_exact_with_reference(gpu_lib, case):
    s, _, _ = make_iface(case, batch=2)
    x = np.stack([np.array(case["x"]), -2.0 * np.array(case["x"])])
    yc = np.stack([np.array(case["yc"]), 3.0 * np.array(case["yc"])]).reshape(2, case["m"])
    Ax = s.spmv(capi.MAT_A, x)
    ATy = s.spmv(capi.MAT_A, yc, transpose=True)
    Hx = s.spmv(capi.MAT_H, x)
    assert Ax[0].tolist() == case["Ax"]
    assert ATy[0].tolist() == case["ATy"]
    assert Hx[0].tolist() == case["Hx"]
    n, m = case["n"], case["m"]
    assert Ax[1].tolist() == orc.csc_times(m, n + 2 * m, case["A_colptr"], case["A_rowidx"], case["A_val"], x[1]).tolist()
    s.close()


def test_batched_segmented_assembly(gpu_lib):
    """Many matrices of ragged sizes in one launch, incl. empty ones and one larger than the shared-memory sort."""
    rng = np.random.default_rng(11)
    mats, seg, ncols, rows, cols = [], [0], [], [], []
    sizes = [(0, 3, 0), (1, 1, 1), (7, 5, 20), (40, 33, 400), (300, 200, 6000), (3, 2, 5)]
    for (nr, nc, z) in sizes:
        if z:
            flat = rng.choice(nr * nc, size=z, replace=False)
            rr, cc = (flat // nc + 1).astype(np.int32), (flat % nc + 1).astype(np.int32)
        else:
            rr, cc = np.zeros(0, np.int32), np.zeros(0, np.int32)
        mats.append((nr, nc, rr, cc))
        rows.append(rr); cols.append(cc); ncols.append(nc); seg.append(seg[-1] + z)
    seg, ncols = np.array(seg, np.int32), np.array(ncols, np.int32)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    colptr = np.zeros(int((ncols + 1).sum()), np.int32)
    rowidx, order = np.zeros(len(rows), np.int32), np.zeros(len(rows), np.int32)
    ms = C.c_float(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = gpu_lib.sqpb200_assemble_csc_batched(0, len(mats), p(seg), p(ncols), p(rows), p(cols), p(colptr), p(rowidx),
                                              p(order), C.byref(ms))
    assert rc == 0 and ms.value > 0
    off = 0
    for k, (nr, nc, rr, cc) in enumerate(mats):
        cp, ri, _, od = orc.csc_from_entries(nc, rr, cc, np.zeros(len(rr)))
        assert colptr[off:off + nc + 1].tolist() == cp.tolist()
        assert rowidx[seg[k]:seg[k + 1]].tolist() == ri.tolist()
        assert order[seg[k]:seg[k + 1]].tolist() == od.tolist()
        off += nc + 1


def test_qphandler_data_kernels(gpu_lib):
    rng = np.random.default_rng(3)
    n, m, B = 5, 3, 7
    hnd = r.QPhandler(r.NLPInfo(nCon=m, nVar=n), r.QPType.QP, batch=B)
    x_l, x_u = -np.abs(rng.standard_normal((B, n))) - 0.1, np.abs(rng.standard_normal((B, n))) + 0.1
    x_k = 0.3 * rng.standard_normal((B, n))
    c_l, c_u, c_k = -np.abs(rng.standard_normal((B, m))), np.abs(rng.standard_normal((B, m))), rng.standard_normal((B, m))
    x_u[0, 0] = 1e18
    c_l[1, 1] = -1e18
    delta = np.abs(rng.standard_normal(B)) + 0.2
    grad, rho = rng.standard_normal((B, n)), 10.0 ** rng.integers(0, 4, B)
    hnd.set_bounds(delta, x_l, x_u, x_k, c_l, c_u, c_k)
    hnd.set_g(grad, rho)
    si = hnd.solverInterface_
    lb, ub, lbA, ubA, g = si.getLb(), si.getUb(), si.getLbA(), si.getUbA(), si.getG()
    for b in range(B):
        elb, eub, elbA, eubA = np.zeros(n + 2 * m), np.zeros(n + 2 * m), np.zeros(m), np.zeros(m)
        orc.qp_bounds(0, n, m, delta[b], x_l[b], x_u[b], x_k[b], c_l[b], c_u[b], c_k[b], elb, eub, elbA, eubA)
        assert (lb[b] == elb).all() and (ub[b] == eub).all() and (lbA[b] == elbA).all() and (ubA[b] == eubA).all()
        assert (g[b] == np.concatenate([grad[b], np.full(2 * m, rho[b])])).all()
    # update_bounds: lbA refreshed, ubA stale (quirk 2); update_delta; update_penalty; update_grad
    hnd.update_bounds(0.5 * delta, x_l, x_u, x_k + 0.1, c_l, c_u, c_k + 1.0)
    hnd.update_penalty(rho * 10)
    lb2, lbA2, ubA2, g2 = si.getLb(), si.getLbA(), si.getUbA(), si.getG()
    assert (ubA2 == ubA).all() and (lbA2 == c_l - (c_k + 1.0)).all()
    assert (lb2[:, :n] == np.maximum(x_l - (x_k + 0.1), -0.5 * delta[:, None])).all()
    assert (g2[:, n:] == (rho * 10)[:, None]).all() and (g2[:, :n] == grad).all()
    hnd.solverInterface_.close()


def test_scalar_setters_and_reset(gpu_lib):
    s = r.CudaQPInterface(nV=4, nC=2, batch=3)
    s.set_lb(2, np.array([1.0, 2.0, 3.0]))
    s.set_ub(0, 7.0)
    s.set_lbA(1, -4.0)
    assert s.getLb()[:, 2].tolist() == [1.0, 2.0, 3.0] and s.getUb()[:, 0].tolist() == [7.0] * 3
    assert s.getLbA()[:, 1].tolist() == [-4.0] * 3
    s.reset_constraints()
    assert not s.getLb().any() and not s.getLbA().any()
    s.close()


def test_batched_assembly_many_small_matrices_warp_kernel(gpu_lib):
    """Many small ragged matrices (padded size <= 512): the one-warp-per-matrix kernel; same index arrays as the reference."""
    rng = np.random.default_rng(12)
    mats, seg, ncols, rows, cols = [], [0], [], [], []
    for k in range(61):
        nr, nc = int(rng.integers(1, 30)), int(rng.integers(1, 40))
        z = int(rng.integers(0, min(nr * nc, 400) + 1)) if k % 7 else 0
        flat = rng.choice(nr * nc, size=z, replace=False) if z else np.zeros(0, np.int64)
        rr, cc = (flat // nc + 1).astype(np.int32), (flat % nc + 1).astype(np.int32)
        mats.append((nr, nc, rr, cc))
        rows.append(rr); cols.append(cc); ncols.append(nc); seg.append(seg[-1] + z)
    seg, ncols = np.array(seg, np.int32), np.array(ncols, np.int32)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    colptr = np.zeros(int((ncols + 1).sum()), np.int32)
    rowidx, order = np.zeros(max(len(rows), 1), np.int32), np.zeros(max(len(rows), 1), np.int32)
    ms = C.c_float(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert gpu_lib.sqpb200_assemble_csc_batched(0, len(mats), p(seg), p(ncols), p(rows), p(cols), p(colptr), p(rowidx), p(order), C.byref(ms)) == 0
    off = 0
    for k, (nr, nc, rr, cc) in enumerate(mats):
        cp, ri, _, od = orc.csc_from_entries(nc, rr, cc, np.zeros(len(rr)))
        assert colptr[off:off + nc + 1].tolist() == cp.tolist()
        assert rowidx[seg[k]:seg[k + 1]].tolist() == ri.tolist()
        assert order[seg[k]:seg[k + 1]].tolist() == od.tolist()
        off += nc + 1


def test_assembly_with_duplicate_triplets(gpu_lib):
    """SURVEY.md 8a quirk 4 (collisions): with duplicate (row, col) triplets ColIndex / RowIndex are those of the reference under
    any tie-break; `order` is the stable one (ties by the original counter), as the oracle and the golden vectors have it."""
    rng = np.random.default_rng(40)
    nr, nc, z = 3, 3, 40
    rr, cc = rng.integers(1, nr + 1, z).astype(np.int32), rng.integers(1, nc + 1, z).astype(np.int32)
    mats = [(nr, nc, rr, cc)] * 9  # nine copies: exercises the one-warp-per-matrix kernel as well
    seg = (np.arange(10) * z).astype(np.int32); ncols = np.full(9, nc, np.int32)
    rows, cols = np.tile(rr, 9), np.tile(cc, 9)
    colptr, rowidx, order = np.zeros(9 * (nc + 1), np.int32), np.zeros(9 * z, np.int32), np.zeros(9 * z, np.int32)
    ms = C.c_float(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert gpu_lib.sqpb200_assemble_csc_batched(0, 9, p(seg), p(ncols), p(rows), p(cols), p(colptr), p(rowidx), p(order), C.byref(ms)) == 0
    cp, ri, _, od = orc.csc_from_entries(nc, rr, cc, np.zeros(z))
    for k in range(9):
        assert colptr[k * (nc + 1):(k + 1) * (nc + 1)].tolist() == cp.tolist()
        assert rowidx[k * z:(k + 1) * z].tolist() == ri.tolist()
        assert order[k * z:(k + 1) * z].tolist() == od.tolist()
    # single-matrix path (one CTA per matrix)
    assert gpu_lib.sqpb200_assemble_csc_batched(0, 1, p(seg[:2].copy()), p(ncols[:1].copy()), p(rr), p(cc), p(colptr), p(rowidx), p(order), C.byref(ms)) == 0
    assert colptr[:nc + 1].tolist() == cp.tolist() and rowidx[:z].tolist() == ri.tolist() and order[:z].tolist() == od.tolist()


def test_randomised_l0_stress(gpu_lib):
    """tools/l0_stress.py at a small size: TMA-staged SpMV / SpMTV and KKT kernels on random shapes and batch sizes (odd tail
    groups, nC = 0, batch 1), bitwise against the oracle's restatement of SpHbMat::times and test_optimality."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "l0_stress.py"), "25", "11"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert " 0 mismatches" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("n", [1, 7, 32, 69, 200])
def test_batched_vector_ops_bit_exact(gpu_lib, n):
    """Row A2: batched Vector::getOneNorm / getInfNorm / times and the in-place elementwise operations (src/Vector.cpp:94-135,
    174-184, 237-251) through the C ABI; sums in the reference's index order -> bit-identical with a sequential evaluation."""
    import ctypes as C
    rng = np.random.default_rng(n)
    B = 1000
    x = rng.standard_normal((B, n)) * 10.0 ** rng.integers(-3, 4, size=(B, n))
    y = rng.standard_normal((B, n))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    out = np.empty(B)
    # sequential left-to-right sums (np.cumsum adds in index order)
    assert gpu_lib.sqpb200_vector_reduce(0, 0, B, n, p(x), None, p(out), capi.LOC_HOST, None) == 0
    assert np.array_equal(out, np.cumsum(np.abs(x), axis=1)[:, -1])
    assert gpu_lib.sqpb200_vector_reduce(0, 1, B, n, p(x), None, p(out), capi.LOC_HOST, None) == 0
    assert np.array_equal(out, np.abs(x).max(axis=1))
    assert gpu_lib.sqpb200_vector_reduce(0, 2, B, n, p(x), p(y), p(out), capi.LOC_HOST, None) == 0
    assert np.array_equal(out, np.cumsum(x * y, axis=1)[:, -1])
    for op, want in ((0, x + y), (1, x - y), (2, y - x), (3, x + 0.25), (4, y), (5, x * 0.25)):
        z = x.copy()
        assert gpu_lib.sqpb200_vector_elementwise(0, op, B, n, p(z), p(y), 0.25, capi.LOC_HOST, None) == 0
        assert np.array_equal(z, want), op
