#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from /root/reference.

Run in the dev container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes
  qp_fixtures.json  the 18 dumped QPs of test/unsolved_QP_data/*.log (QORE layout, converted to the
                    qpOASES layout exactly as test/QPsolvers_testers.cpp:18-29,206-218 does: CSR ->
                    dense -> CSC keeping |v| > m_eps, bounds split at nV) and the 10 embedded QPs of
                    test/unsolved_QPs/*.hpp (already CSC).  These are INPUTS only: the reference
                    stores no expected outputs.
  l0_golden.json    outputs of the reference's own L0 code (oracle/_ref/libref_l0.so): CSC index
                    arrays, `order`, refreshed values, SpMV / SpMTV results and norms for the HS071
                    shaped probe of SURVEY.md section 8c and for seeded random matrices.
  qp_fixtures_qore.json  the 18 `.log` dumps unconverted (QORE layout, explicit zeros kept): inputs of the QORE data constructor.
  qore_golden.json  the same triplets through the reference's compressed-row SpHbMat (the QORE layout of
                    src/QOREInterface.cpp:89-90, 643-659): row pointers, column indices, `order`, values.
"""
import ctypes as C
import glob
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
M_EPS = 1.0e-16
INFTY = 1.0e20


def csr_to_csc_via_dense(nrow, ncol, rowptr, colidx, val):
    """test/QPsolvers_testers.cpp:18-29 -> SpHbMat::get_dense_matrix + dense constructor
    (src/SpHbMat.cpp:59-165, column-compressed, row_oriented=true branch)."""
    dense = np.zeros((nrow, ncol))
    for r in range(nrow):
        for e in range(rowptr[r], rowptr[r + 1]):
            dense[r, colidx[e]] = val[e]
    colptr, rowidx, vals = [0], [], []
    for j in range(ncol):
        for i in range(nrow):
            if abs(dense[i, j]) > M_EPS:
                rowidx.append(i)
                vals.append(float(dense[i, j]))
        colptr.append(len(rowidx))
    return colptr, rowidx, vals


def read_log(path):
    tok = open(path).read().split()
    it = iter(tok)
    nV, nC, zA, zH = (int(next(it)) for _ in range(4))
    lbq = [float(next(it)) for _ in range(nV + nC)]
    ubq = [float(next(it)) for _ in range(nV + nC)]
    g = [float(next(it)) for _ in range(nV)]
    A_rp = [int(next(it)) for _ in range(nC + 1)]
    A_ci = [int(next(it)) for _ in range(zA)]
    A_v = [float(next(it)) for _ in range(zA)]
    H_rp = [int(next(it)) for _ in range(nV + 1)]
    H_ci = [int(next(it)) for _ in range(zH)]
    H_v = [float(next(it)) for _ in range(zH)]
    Ap, Ai, Av = csr_to_csc_via_dense(nC, nV, A_rp, A_ci, A_v)
    Hp, Hi, Hv = csr_to_csc_via_dense(nV, nV, H_rp, H_ci, H_v)
    return dict(name=os.path.basename(path).replace("qpdata.log", ""), source="log", nV=nV, nC=nC,
                lb=lbq[:nV], ub=ubq[:nV], lbA=lbq[nV:], ubA=ubq[nV:], g=g,
                A_colptr=Ap, A_rowidx=Ai, A_val=Av, H_colptr=Hp, H_rowidx=Hi, H_val=Hv)


def read_hpp(path):
    txt = open(path).read()
    arrs = {}
    for m in re.finditer(r"(real_t|sparse_int_t)\s+(\w+)\[\]\s*=\s*\{([^}]*)\}", txt):
        kind, name, body = m.groups()
        vals = []
        for t in body.replace("\n", " ").split(","):
            t = t.strip()
            if not t:
                continue
            if kind == "sparse_int_t":
                vals.append(int(t))
            else:
                v = float(t)  # handles inf / -inf
                vals.append(max(-INFTY, min(INFTY, v)))
        arrs[name] = vals
    nV, nC = len(arrs["lb"]), len(arrs["lbA"])
    assert len(arrs["A_jc"]) == nV + 1 and len(arrs["H_jc"]) == nV + 1
    return dict(name=os.path.basename(path).replace(".hpp", "") + "_hpp", source="hpp", nV=nV, nC=nC,
                lb=arrs["lb"], ub=arrs["ub"], lbA=arrs["lbA"], ubA=arrs["ubA"], g=arrs["g"],
                A_colptr=arrs["A_jc"], A_rowidx=arrs["A_ir"], A_val=arrs["A_val"],
                H_colptr=arrs["H_jc"], H_rowidx=arrs["H_ir"], H_val=arrs["H_val"])


def make_qp_fixtures():
    qps = [read_log(p) for p in sorted(glob.glob(REF + "/test/unsolved_QP_data/*.log"))]
    qps += [read_hpp(p) for p in sorted(glob.glob(REF + "/test/unsolved_QPs/*.hpp"))]
    with open(os.path.join(OUT, "qp_fixtures.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", qps=qps), f)
    print("qp_fixtures.json:", len(qps), "QPs;", [(q["name"], q["nV"], q["nC"]) for q in qps])


# ------------------------------------------------------------------ L0 golden from the reference code
ip, dp = orc._ip, orc._dp


def ref_A(R, nrow, ncol, row1, col1, val, iinfo, val2):
    row1, col1, val, val2 = orc._i32(row1), orc._i32(col1), orc._f64(val), orc._f64(val2)
    irow, jcol, size, ival = [a.copy() for a in iinfo]
    z = len(row1) + int(size.sum())
    colptr, rowidx, order = np.zeros(ncol + 1, np.int32), np.zeros(z, np.int32), np.zeros(z, np.int32)
    v0, v1 = np.zeros(z), np.zeros(z)
    R.ref_assemble_A(nrow, ncol, len(row1), ip(row1), ip(col1), dp(val), len(size), ip(irow), ip(jcol), ip(size),
                     dp(ival), None, ip(colptr), ip(rowidx), dp(v0), ip(order))
    R.ref_assemble_A(nrow, ncol, len(row1), ip(row1), ip(col1), dp(val), len(size), ip(irow), ip(jcol), ip(size),
                     dp(ival), dp(val2), ip(colptr), ip(rowidx), dp(v1), ip(order))
    return colptr, rowidx, order, v0, v1


def ref_H(R, n, row1, col1, val, val2):
    row1, col1, val, val2 = orc._i32(row1), orc._i32(col1), orc._f64(val), orc._f64(val2)
    zmax = 2 * len(row1)
    colptr, rowidx, order = np.zeros(n + 1, np.int32), np.zeros(zmax, np.int32), np.zeros(zmax, np.int32)
    v0, v1 = np.zeros(zmax), np.zeros(zmax)
    z = R.ref_assemble_H(n, len(row1), ip(row1), ip(col1), dp(val), 1, None, ip(colptr), ip(rowidx), dp(v0), ip(order))
    R.ref_assemble_H(n, len(row1), ip(row1), ip(col1), dp(val), 1, dp(val2), ip(colptr), ip(rowidx), dp(v1), ip(order))
    return colptr, rowidx[:z], order[:z], v0[:z], v1[:z]


def ref_times(R, nrow, ncol, colptr, rowidx, val, x, transpose):
    y = np.zeros(ncol if transpose else nrow)
    f = R.ref_csc_transposed_times if transpose else R.ref_csc_times
    f(nrow, ncol, len(rowidx), ip(orc._i32(colptr)), ip(orc._i32(rowidx)), dp(orc._f64(val)), dp(orc._f64(x)), dp(y))
    return y


def random_case(rng, n, m, dens_j, dens_h):
    """Unique-key triplets in column-major order (the order AmplTNLP emits)."""
    jr, jc = [], []
    for c in range(n):
        for r in range(m):
            if rng.random() < dens_j:
                jr.append(r + 1); jc.append(c + 1)
    hr, hc = [], []
    for c in range(n):
        for r in range(c + 1):  # upper triangle, column by column
            if rng.random() < dens_h or r == c:
                hr.append(r + 1); hc.append(c + 1)
    return jr, jc, hr, hc


def make_l0_golden():
    R = orc.ref_lib()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    cases = []
    rng = np.random.default_rng(20261018)
    specs = [("hs071", 4, 2, None), ("hs071_rowmajor", 4, 2, -1)] + [("rand%d" % i, int(rng.integers(1, 24)), int(rng.integers(0, 30)), i)
                                         for i in range(12)]
    for name, n, m, seed in specs:
        if seed is None:  # dense 2x4 Jacobian, dense upper-triangular Hessian (hs071.nl)
            jr = [1, 2] * 4; jc = [1, 1, 2, 2, 3, 3, 4, 4]
            hr = [1, 1, 2, 1, 2, 3, 1, 2, 3, 4]; hc = [1, 2, 2, 3, 3, 3, 4, 4, 4, 4]
        elif seed == -1:  # the row-major probe of SURVEY.md section 8c (lower-triangular Hessian)
            jr = [1] * 4 + [2] * 4; jc = [1, 2, 3, 4] * 2
            hr = [1, 2, 2, 3, 3, 3, 4, 4, 4, 4]; hc = [1, 1, 2, 1, 2, 3, 1, 2, 3, 4]
        else:
            jr, jc, hr, hc = random_case(rng, n, m, 0.4, 0.3)
            if seed % 2 == 1:  # also exercise a non-trivial `order`: shuffled triplets (keys stay unique)
                pj, ph = rng.permutation(len(jr)), rng.permutation(len(hr))
                jr, jc = [jr[i] for i in pj], [jc[i] for i in pj]
                hr, hc = [hr[i] for i in ph], [hc[i] for i in ph]
                # store some Hessian entries in the lower triangle instead
                for i in range(0, len(hr), 3):
                    hr[i], hc[i] = hc[i], hr[i]
        nV, nC = n + 2 * m, m
        jv, jv2 = rng.standard_normal(len(jr)), rng.standard_normal(len(jr))
        hv, hv2 = rng.standard_normal(len(hr)), rng.standard_normal(len(hr))
        iinfo = orc.identity_info(n, m)
        Ap, Ai, Ao, Av0, Av1 = ref_A(R, nC, nV, jr, jc, jv, iinfo, jv2)
        Hp, Hi, Ho, Hv0, Hv1 = ref_H(R, nV, hr, hc, hv, hv2)
        x, yc = rng.standard_normal(nV), rng.standard_normal(nC)
        case = dict(name=name, n=n, m=m, J_row1=jr, J_col1=jc, J_val=jv.tolist(), J_val2=jv2.tolist(),
                    H_row1=hr, H_col1=hc, H_val=hv.tolist(), H_val2=hv2.tolist(),
                    A_colptr=Ap.tolist(), A_rowidx=Ai.tolist(), A_order=Ao.tolist(), A_val=Av0.tolist(),
                    A_val2=Av1.tolist(), H_colptr=Hp.tolist(), H_rowidx=Hi.tolist(), H_order=Ho.tolist(),
                    H_cscval=Hv0.tolist(), H_cscval2=Hv1.tolist(), x=x.tolist(), yc=yc.tolist(),
                    Ax=ref_times(R, nC, nV, Ap, Ai, Av0, x, False).tolist(),
                    ATy=ref_times(R, nC, nV, Ap, Ai, Av0, yc, True).tolist(),
                    Hx=ref_times(R, nV, nV, Hp, Hi, Hv0, x, False).tolist(),
                    one_norm_x=R.ref_one_norm(dp(x), nV), inf_norm_x=R.ref_inf_norm(dp(x), nV))
        cases.append(case)
    with open(os.path.join(OUT, "l0_golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py (reference L0 via oracle/_ref/libref_l0.so)",
                       cases=cases), f)
    print("l0_golden.json:", len(cases), "cases; hs071 A.colptr =", cases[0]["A_colptr"], "A.order =",
          cases[0]["A_order"], "H.colptr =", cases[0]["H_colptr"])


def make_qore_raw_fixtures():
    """qp_fixtures_qore.json: the 18 `.log` dumps as they are (QORE layout: stacked bounds, row-compressed A and H with their
    explicit zeros), i.e. what test/QPsolvers_testers.cpp:48-150 reads and hands to the QORE data constructor (:74-75, 172-175)
    before it converts a copy for qpOASES.  Inputs only."""
    qps = []
    for p in sorted(glob.glob(REF + "/test/unsolved_QP_data/*.log")):
        it = iter(open(p).read().split())
        nV, nC, zA, zH = (int(next(it)) for _ in range(4))
        f = lambda n: [float(next(it)) for _ in range(n)]
        i = lambda n: [int(next(it)) for _ in range(n)]
        lb, ub, g = f(nV + nC), f(nV + nC), f(nV)
        A_rp, A_ci, A_v = i(nC + 1), i(zA), f(zA)
        H_rp, H_ci, H_v = i(nV + 1), i(zH), f(zH)
        qps.append(dict(name=os.path.basename(p).replace("qpdata.log", ""), nV=nV, nC=nC, lb=lb, ub=ub, g=g, A_rowptr=A_rp,
                        A_colidx=A_ci, A_val=A_v, H_rowptr=H_rp, H_colidx=H_ci, H_val=H_v))
    with open(os.path.join(OUT, "qp_fixtures_qore.json"), "w") as fo:
        json.dump(dict(generator="tests/golden/make_golden.py (test/unsolved_QP_data/*.log, unconverted)", qps=qps), fo)
    print("qp_fixtures_qore.json:", len(qps), "QPs; explicit zeros in H:", {q["name"]: sum(1 for v in q["H_val"] if v == 0.0) for q in qps})


def make_qore_golden():
    """qore_golden.json: the compressed-row arrays the reference's QOREInterface hands to QPSetData
    (src/QOREInterface.cpp:89-90): SpHbMat(..., isCompressedRow = true)::setStructure / setMatVal run on the triplets of
    every l0_golden.json case (row pointers, column indices, `order`, values before and after a refresh) and
    SpHbMat::times on the compressed-row matrix."""
    R = orc.ref_lib()
    assert R is not None and hasattr(R, "ref_assemble_A_csr"), "build oracle/_ref first (make -C oracle ref)"
    src = json.load(open(os.path.join(OUT, "l0_golden.json")))["cases"]
    cases = []
    for c in src:
        n, m = c["n"], c["m"]
        nV, nC = n + 2 * m, m
        jr, jc, jv, jv2 = orc._i32(c["J_row1"]), orc._i32(c["J_col1"]), orc._f64(c["J_val"]), orc._f64(c["J_val2"])
        hr, hc, hv, hv2 = orc._i32(c["H_row1"]), orc._i32(c["H_col1"]), orc._f64(c["H_val"]), orc._f64(c["H_val2"])
        irow, jcol, size, ival = [a.copy() for a in orc.identity_info(n, m)]
        z = len(jr) + 2 * m
        Arp, Aci, Ao = np.zeros(nC + 1, np.int32), np.zeros(z, np.int32), np.zeros(z, np.int32)
        Av0, Av1 = np.zeros(z), np.zeros(z)
        for vref, out in ((None, Av0), (dp(jv2), Av1)):
            R.ref_assemble_A_csr(nC, nV, len(jr), ip(jr), ip(jc), dp(jv), 2, ip(irow), ip(jcol), ip(size), dp(ival), vref,
                                 ip(Arp), ip(Aci), dp(out), ip(Ao))
        zmax = 2 * len(hr)
        Hrp, Hci, Ho = np.zeros(nV + 1, np.int32), np.zeros(zmax, np.int32), np.zeros(zmax, np.int32)
        Hv0, Hv1 = np.zeros(zmax), np.zeros(zmax)
        for vref, out in ((None, Hv0), (dp(hv2), Hv1)):
            zh = R.ref_assemble_H_csr(nV, len(hr), ip(hr), ip(hc), dp(hv), 1, vref, ip(Hrp), ip(Hci), dp(out), ip(Ho))
        x = orc._f64(c["x"])
        Ax, Hx = np.zeros(nC), np.zeros(nV)
        R.ref_csr_times(nC, nV, z, ip(Arp), ip(Aci), dp(Av0), dp(x), dp(Ax))
        R.ref_csr_times(nV, nV, zh, ip(Hrp), ip(Hci[:zh].copy()), dp(Hv0[:zh].copy()), dp(x), dp(Hx))
        cases.append(dict(name=c["name"], n=n, m=m, A_rowptr=Arp.tolist(), A_colidx=Aci.tolist(), A_order=Ao.tolist(),
                          A_val=Av0.tolist(), A_val2=Av1.tolist(), H_rowptr=Hrp.tolist(), H_colidx=Hci[:zh].tolist(),
                          H_order=Ho[:zh].tolist(), H_val=Hv0[:zh].tolist(), H_val2=Hv1[:zh].tolist(), Ax=Ax.tolist(),
                          Hx=Hx.tolist()))
    with open(os.path.join(OUT, "qore_golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py (reference compressed-row SpHbMat via oracle/_ref/libref_l0.so; "
                                 "inputs = the triplets of l0_golden.json)", cases=cases), f)
    print("qore_golden.json:", len(cases), "cases; hs071 A.rowptr =", cases[0]["A_rowptr"], "A.order =", cases[0]["A_order"])


if __name__ == "__main__":
    orc.build()
    make_qp_fixtures()
    make_l0_golden()
    make_qore_golden()
    make_qore_raw_fixtures()
