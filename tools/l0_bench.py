"""HBM roofline of the L0 kernels: batched SpMV / SpMTV (A7), value scatter (A6), QP data construction (B2/B3),
stand-alone KKT test (C8) and segmented triplet -> CSC assembly (A4/A5).  Inputs and outputs stay in device memory
(loc = DEVICE); CUDA events around `reps` back-to-back launches; algorithmic bytes per unit as in SURVEY 8d; peak =
MEASURED_PEAKS.json hbm_gbs.  Working sets are far larger than the 126 MB L2."""
import sys, os, json, ctypes as C, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200 import capi
import helpers as H
L = capi.lib()
try: PEAK = json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: PEAK = 6650.0
dev = torch.device("cuda", 0)

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def row(name, bytes_, ms):
    gbs = bytes_ / (ms * 1e-3) / 1e9
    print(f"| {name} | {bytes_/1e6:.1f} | {ms:.3f} | {gbs:.0f} | {gbs/PEAK:.2f} |", flush=True)

def shape_case(label, nV, nC, Ac, Hc, n, m, B):
    zA, zH = len(Ac[1]), len(Hc[1])
    print(f"\n### {label}: nV={nV}, nC={nC}, nnz(A)={zA}, nnz(H)={zH}, batch {B}\n\n| kernel | algorithmic MB / launch | ms | GB/s | frac of {PEAK:.0f} GB/s |\n|---|---|---|---|---|")
    s = r.CudaQPInterface(nV=nV, nC=nC, qptype=r.QPType.QP, batch=B, keep_state=False)
    g = torch.Generator(device=dev); g.manual_seed(1)
    rnd = lambda *sh: torch.randn(*sh, dtype=torch.float64, device=dev, generator=g)
    s.set_csc(capi.MAT_A, Ac[0], Ac[1], rnd(B, zA)); s.set_csc(capi.MAT_H, Hc[0], Hc[1], rnd(B, zH))
    x, yc = rnd(B, nV), rnd(B, nC)
    oC, oV = torch.empty(B, nC, dtype=torch.float64, device=dev), torch.empty(B, nV, dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    row("`spmv_kernel` A x", 8 * (zA + nV + nC) * B, timed(lambda: L.sqpb200_spmv(s.h, capi.MAT_A, 0, p(x), p(oC), capi.LOC_DEVICE)))
    row("`spmv_kernel` A'y", 8 * (zA + nV + nC) * B, timed(lambda: L.sqpb200_spmv(s.h, capi.MAT_A, 1, p(yc), p(oV), capi.LOC_DEVICE)))
    row("`spmv_kernel` H x", 8 * (zH + 2 * nV) * B, timed(lambda: L.sqpb200_spmv(s.h, capi.MAT_H, 0, p(x), p(oV), capi.LOC_DEVICE)))
    av, hv = rnd(B, zA), rnd(B, zH)
    row("value refresh A (CSC order, DMA)", 16 * zA * B, timed(lambda: L.sqpb200_set_values_csc(s.h, capi.MAT_A, p(av), capi.LOC_DEVICE, 0)))
    # QP data construction
    delta = torch.ones(B, dtype=torch.float64, device=dev); rho = torch.ones(B, dtype=torch.float64, device=dev)
    xl, xu, xk, cl, cu, ck, gr = rnd(B, n), rnd(B, n), rnd(B, n), rnd(B, m), rnd(B, m), rnd(B, m), rnd(B, n)
    row("`qphandler_bounds_kernel` (set_bounds)", 8 * (3 * n + 3 * m + 1 + 2 * nV + 2 * nC) * B,
        timed(lambda: L.sqpb200_qphandler_bounds(s.h, 0, n, m, p(delta), p(xl), p(xu), p(xk), p(cl), p(cu), p(ck), capi.LOC_DEVICE)))
    row("`qphandler_g_kernel`", 8 * (n + 1 + nV) * B, timed(lambda: L.sqpb200_qphandler_g(s.h, n, m, p(gr), p(rho), capi.LOC_DEVICE)))
    # KKT test on whatever x, y the handle holds (zeros): same traffic
    out = torch.empty(B, 5, dtype=torch.float64, device=dev)
    row("`kkt_kernel` (test_optimality)", (8 * (zA + zH + 5 * nV + 2 * nC + (nV + nC)) + 2 * (nV + nC) + 4 * (nV + nC) + 40) * B,
        timed(lambda: L.sqpb200_kkt_residuals_recompute(s.h, p(out), capi.LOC_DEVICE)))
    s.close()

def scatter_case(label, n, m, jr, jc, B):
    info = r.NLPInfo(nCon=m, nVar=n, nnz_jac_g=len(jr), nnz_h_lag=0)
    s = r.CudaQPInterface(info, r.QPType.LP, batch=B, keep_state=False)
    I = r.IdentityInfo(irow=np.array([1, 1], np.int32), jcol=np.array([n + 1, n + m + 1], np.int32), size=np.array([m, m], np.int32), value=np.array([1.0, -1.0]))
    vals = torch.randn(B, len(jr), dtype=torch.float64, device=dev)
    s.set_A(r.SpTripletMat(np.array(jr, np.int32), np.array(jc, np.int32), vals, m, n, False), I)
    zJ, zA = len(jr), len(jr) + 2 * m
    ms = timed(lambda: L.sqpb200_set_values_A(s.h, C.c_void_p(vals.data_ptr()), capi.LOC_DEVICE, 0))
    print(f"\n### value scatter (A6), {label}: zJ={zJ}, batch {B}\n\n| kernel | algorithmic MB / launch | ms | GB/s | frac |\n|---|---|---|---|---|")
    row("`scatter_values_kernel` (setMatVal)", 20 * zJ * B, ms)
    s.close()

def assembly_case(nmat, n, m, jr, jc):
    z = len(jr) + 2 * m
    er = np.concatenate([jr, 1 + np.arange(m), 1 + np.arange(m)]).astype(np.int32); ec = np.concatenate([jc, n + 1 + np.arange(m), n + m + 1 + np.arange(m)]).astype(np.int32)
    row1, col1 = np.tile(er, nmat), np.tile(ec, nmat)
    seg = (np.arange(nmat + 1) * z).astype(np.int32); ncol = np.full(nmat, n + 2 * m, np.int32)
    colptr, rowidx, order = np.zeros(nmat * (n + 2 * m + 1), np.int32), np.zeros(nmat * z, np.int32), np.zeros(nmat * z, np.int32)
    ms = C.c_float(0)
    ip = lambda a: a.ctypes.data_as(C.c_void_p)
    for _ in range(2):
        rc = L.sqpb200_assemble_csc_batched(0, nmat, ip(seg), ip(ncol), ip(row1), ip(col1), ip(colptr), ip(rowidx), ip(order), C.byref(ms))
    assert rc == 0
    print(f"\n### segmented triplet -> CSC (A4/A5): {nmat} matrices of z={z} entries, {n + 2 * m} columns, one launch\n\n| kernel | algorithmic MB / launch | ms | GB/s | frac |\n|---|---|---|---|---|")
    row("`csc_assemble_kernel`", (28 * z + 4 * (n + 2 * m + 1)) * nmat, ms.value)

fx = {q["name"]: q for q in H.load_qp_fixtures()}
q = fx["QORE_hs116"]
shape_case("hs116 shape (largest dumped QP)", q["nV"], q["nC"], (q["A_colptr"], q["A_rowidx"]), (q["H_colptr"], q["H_rowidx"]), 13, 28, 1 << 19)
d = H.synthetic_large_qp(256)
shape_case("config-4 n=256 shape", d["nV"], d["nC"], d["Ac"][:2], d["Hc"][:2], 256, 128, 1 << 15)
# Jacobian triplets of the hs116 shape: the structural part of A without the identity columns, column-major
A = q; jr, jc = [], []
for c in range(13):
    for e in range(A["A_colptr"][c], A["A_colptr"][c + 1]):
        jr.append(A["A_rowidx"][e] + 1); jc.append(c + 1)
scatter_case("hs116 shape", 13, 28, jr, jc, 1 << 20)
assembly_case(1 << 16, 13, 28, np.array(jr), np.array(jc))
