"""QP dump formats of the reference and the replay constructor.

* `.log`  QORE layout written by QOREInterface::WriteQPDataToFile (src/QOREInterface.cpp:582-598) and read back by the
  replay driver (test/QPsolvers_testers.cpp:48-150): one number per line -- nV, nC, nnz(A), nnz(H); lb[nV+nC] and
  ub[nV+nC] (variable bounds first, then constraint bounds); g[nV]; A as CSR (rowptr[nC+1], colidx, val); H as CSR.
* `.hpp`  qpOASES C arrays (test/unsolved_QPs/*.hpp): `real_t lb[] = {...}`, `sparse_int_t A_jc[] = {...}` etc., CSC.

`read_*` return the same dict the fixtures use (CSC `A_colptr/A_rowidx/A_val`, `H_*`, g, lb, ub, lbA, ubA) with the CSR -> CSC
conversion of the replay driver (through the dense matrix, entries with |v| <= 1e-16 dropped: src/SpHbMat.cpp:59-165 as
called at test/QPsolvers_testers.cpp:206-218).  `replay` is the data constructor of the plugin family
(include/sqphot/qpOASESInterface.hpp:44-51): a backend built directly from H, A, g, lb, ub, lbA, ubA.
"""
import os
import re

import numpy as np

M_EPS = 1.0e-16
INFTY = 1.0e20


def _csr_to_csc_via_dense(nrow, ncol, rowptr, colidx, val):
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


def _csc_to_csr(nrow, ncol, colptr, rowidx, val):
    rows = [[] for _ in range(nrow)]
    for c in range(ncol):
        for e in range(colptr[c], colptr[c + 1]):
            rows[rowidx[e]].append((c, val[e]))
    rp, ci, v = [0], [], []
    for r in range(nrow):
        for c, x in rows[r]:
            ci.append(c)
            v.append(x)
        rp.append(len(ci))
    return rp, ci, v


def read_qore_log(path):
    it = iter(open(path).read().split())
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
    Ap, Ai, Av = _csr_to_csc_via_dense(nC, nV, A_rp, A_ci, A_v)
    Hp, Hi, Hv = _csr_to_csc_via_dense(nV, nV, H_rp, H_ci, H_v)
    return dict(name=os.path.basename(path).replace("qpdata.log", ""), source="log", nV=nV, nC=nC,
                lb=lbq[:nV], ub=ubq[:nV], lbA=lbq[nV:], ubA=ubq[nV:], g=g,
                A_colptr=Ap, A_rowidx=Ai, A_val=Av, H_colptr=Hp, H_rowidx=Hi, H_val=Hv)


def write_qore_log(path, q):
    """Inverse of read_qore_log: the layout of QOREInterface::WriteQPDataToFile."""
    nV, nC = q["nV"], q["nC"]
    A = _csc_to_csr(nC, nV, q["A_colptr"], q["A_rowidx"], q["A_val"])
    H = _csc_to_csr(nV, nV, q["H_colptr"], q["H_rowidx"], q["H_val"])
    with open(path, "w") as f:
        for k in (nV, nC, len(A[2]), len(H[2])):
            f.write("%d\n" % k)
        for v in (list(q["lb"]) + list(q["lbA"]), list(q["ub"]) + list(q["ubA"]), q["g"]):
            for t in v:
                f.write("%23.16e\n" % t)
        for M in (A, H):
            for t in M[0]:
                f.write("%d\n" % t)
            for t in M[1]:
                f.write("%d\n" % t)
            for t in M[2]:
                f.write("%23.16e\n" % t)


def read_qpoases_hpp(path):
    txt = open(path).read()
    arrs = {}
    for m in re.finditer(r"(real_t|sparse_int_t)\s+(\w+)\[\]\s*=\s*\{([^}]*)\}", txt):
        kind, name, body = m.groups()
        vals = []
        for t in body.replace("\n", " ").split(","):
            t = t.strip()
            if t:
                vals.append(int(t) if kind == "sparse_int_t" else max(-INFTY, min(INFTY, float(t))))
        arrs[name] = vals
    nV, nC = len(arrs["lb"]), len(arrs["lbA"])
    if len(arrs["A_jc"]) != nV + 1 or len(arrs["H_jc"]) != nV + 1:
        raise ValueError("%s: column pointers do not match nV=%d" % (path, nV))
    return dict(name=os.path.basename(path).replace(".hpp", "") + "_hpp", source="hpp", nV=nV, nC=nC,
                lb=arrs["lb"], ub=arrs["ub"], lbA=arrs["lbA"], ubA=arrs["ubA"], g=arrs["g"],
                A_colptr=arrs["A_jc"], A_rowidx=arrs["A_ir"], A_val=arrs["A_val"],
                H_colptr=arrs["H_jc"], H_rowidx=arrs["H_ir"], H_val=arrs["H_val"])


def write_qpoases_hpp(path, q):
    def arr(kind, name, vals, fmt):
        return "%s %s[] = {%s};\n" % (kind, name, ", ".join(fmt % v for v in vals))
    with open(path, "w") as f:
        f.write("// QP dump in the layout of test/unsolved_QPs/*.hpp (qpOASES sparse arrays, CSC)\n")
        for M in ("H", "A"):
            f.write(arr("sparse_int_t", M + "_ir", q[M + "_rowidx"], "%d"))
            f.write(arr("sparse_int_t", M + "_jc", q[M + "_colptr"], "%d"))
            f.write(arr("real_t", M + "_val", q[M + "_val"], "%.16e"))
        for k in ("g", "lb", "ub", "lbA", "ubA"):
            f.write(arr("real_t", k, q[k], "%.16e"))


def read_dump(path):
    return read_qpoases_hpp(path) if path.endswith(".hpp") else read_qore_log(path)


def replay(q, batch=1, device=0, options=None, **kw):
    """Data constructor (include/sqphot/qpOASESInterface.hpp:44-51): a CUDA backend loaded with the dumped QP, every instance
    of the batch holding the same data; the caller perturbs / solves it (optimizeQP)."""
    from . import capi
    from .qp_interface import CudaQPInterface
    from .sqp_types import QPType
    if isinstance(q, str):
        q = read_dump(q)
    nV, nC = q["nV"], q["nC"]
    s = CudaQPInterface(nV=nV, nC=nC, qptype=QPType.QP, options=options, batch=batch, device=device, **kw)
    s.set_csc(capi.MAT_A, q["A_colptr"], q["A_rowidx"], np.asarray(q["A_val"], dtype=np.float64))
    s.set_csc(capi.MAT_H, q["H_colptr"], q["H_rowidx"], np.asarray(q["H_val"], dtype=np.float64))
    tile = lambda k, n: np.ascontiguousarray(np.tile(np.asarray(q[k], dtype=np.float64).reshape(1, n), (batch, 1)))
    s.set_g(tile("g", nV)); s.set_lb(tile("lb", nV)); s.set_ub(tile("ub", nV))
    if nC:
        s.set_lbA(tile("lbA", nC)); s.set_ubA(tile("ubA", nC))
    return s
