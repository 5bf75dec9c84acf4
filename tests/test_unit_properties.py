"""The reference's own unit tests, restated (test/unitTest/test_SpHbMat.cpp, test_SpTripletMat.cpp; helpers
unit_test_utils.hpp:4-28): a random matrix of at most 10 x 10 with integer entries 1..10 (test_SpHbMat.cpp:410-430), then
  * dense -> sparse -> dense round trips in both compressed layouts (:11-80),
  * SpMV and SpMTV against a dense double loop with exact `==` (exact because the entries are small integers, :83-230),
  * the triplet -> Harwell-Boeing -> triplet round trip (:317-380).
The reference seeds from time(NULL) and returns 0 whatever happens; here the seed is fixed, 60 matrices are drawn and every
comparison is an assertion.  CPU: the oracle's restatement, and the reference's own classes when oracle/_ref is built.  GPU: the
library's assembly, value scatter and SpMV / SpMTV kernels through the C ABI, in the column- and the row-compressed layout."""
import numpy as np
import pytest

import restartsqp_b200 as r
from restartsqp_b200 import capi
from oracle import oracle_py as orc


def random_cases(count=60, seed=20261019):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        nr, nc = int(rng.integers(1, 11)), int(rng.integers(1, 11))
        z = int(rng.integers(1, nr * nc + 1))
        D = np.zeros(nr * nc)
        D[rng.permutation(nr * nc)[:z]] = rng.integers(1, 11, z)
        D = D.reshape(nr, nc)
        rr, cc = np.nonzero(D)
        p = rng.permutation(len(rr))  # triplets in arbitrary order
        rr, cc = rr[p], cc[p]
        yield D, (rr + 1).astype(np.int32), (cc + 1).astype(np.int32), D[rr, cc].copy(), rng.integers(1, 11, nc).astype(np.float64), \
            rng.integers(1, 11, nr).astype(np.float64)


def dense_from_csc(nr, nc, colptr, rowidx, val):
    D = np.zeros((nr, nc))
    for c in range(nc):
        for e in range(colptr[c], colptr[c + 1]):
            D[rowidx[e], c] += val[e]
    return D


def dense_from_csr(nr, nc, rowptr, colidx, val):
    D = np.zeros((nr, nc))
    for i in range(nr):
        for e in range(rowptr[i], rowptr[i + 1]):
            D[i, colidx[e]] += val[e]
    return D


def test_oracle_round_trips_and_products_are_exact():
    R = orc.ref_lib()
    ip, dp = orc._ip, orc._dp
    for D, rr, cc, vv, x, y in random_cases():
        nr, nc = D.shape
        cp, ri, v, od = orc.csc_from_entries(nc, rr, cc, vv)
        assert (dense_from_csc(nr, nc, cp, ri, v) == D).all()
        rp, ci, v2, od2 = orc.csr_from_entries(nr, rr, cc, vv)
        assert (dense_from_csr(nr, nc, rp, ci, v2) == D).all()
        # triplet -> HB -> triplet: entry k of the triplet list sits at order[k]
        assert (v[od] == vv).all() and (ri[od] == rr - 1).all() and (v2[od2] == vv).all() and (ci[od2] == cc - 1).all()
        Ax = orc.csc_times(nr, nc, cp, ri, v, x)
        ATy = orc.csc_times(nr, nc, cp, ri, v, y, transpose=True)
        assert (Ax == D @ x).all() and (ATy == D.T @ y).all()
        if R is not None:  # the reference's own SpHbMat / SpTripletMat on the same matrix
            out = np.zeros(nr)
            R.ref_csc_times(nr, nc, len(ri), ip(cp), ip(ri), dp(v), dp(x), dp(out))
            assert (out == D @ x).all()
            if hasattr(R, "ref_csr_times"):
                R.ref_csr_times(nr, nc, len(ci), ip(rp), ip(ci), dp(v2), dp(x), dp(out))
                assert (out == D @ x).all()
            outT = np.zeros(nc)
            R.ref_csc_transposed_times(nr, nc, len(ri), ip(cp), ip(ri), dp(v), dp(y), dp(outT))
            assert (outT == D.T @ y).all()
            R.ref_triplet_times(nr, nc, len(rr), ip(rr), ip(cc), dp(vv), 0, 0, dp(x), dp(out))
            assert (out == D @ x).all()
            R.ref_triplet_times(nr, nc, len(rr), ip(rr), ip(cc), dp(vv), 0, 1, dp(y), dp(outT))
            assert (outT == D.T @ y).all()


@pytest.mark.gpu
def test_device_round_trips_and_products_are_exact(gpu_lib):
    B = 3
    for D, rr, cc, vv, x, y in random_cases(40):
        nr, nc = D.shape
        for layout in ("csc", "csr"):
            s = r.CudaQPInterface(nV=nc, nC=nr, qptype=r.QPType.LP, batch=B)
            T = r.SpTripletMat(rr, cc, np.stack([vv, 2.0 * vv, vv]), nr, nc, False)
            if layout == "csc":
                s.set_A(T, None)
            else:
                s.set_A_csr(T, None)
            A = s.getA()
            assert (dense_from_csc(nr, nc, A["ColIndex"], A["RowIndex"], A["MatVal"][0]) == D).all()
            assert (dense_from_csc(nr, nc, A["ColIndex"], A["RowIndex"], A["MatVal"][1]) == 2.0 * D).all()
            assert (A["MatVal"][2][A["order"]] == vv).all() and (A["RowIndex"][A["order"]] == rr - 1).all()
            if layout == "csr":
                Q = s.get_csr(capi.MAT_A)
                assert (dense_from_csr(nr, nc, Q["RowIndex"], Q["ColIndex"], Q["MatVal"][1]) == 2.0 * D).all()
                assert (Q["MatVal"][0][Q["order"]] == vv).all() and (Q["ColIndex"][Q["order"]] == cc - 1).all()
            X = np.stack([x, 3.0 * x, -x])
            Y = np.stack([y, -2.0 * y, y])
            Ax, ATy = s.spmv(capi.MAT_A, X), s.spmv(capi.MAT_A, Y, transpose=True)
            assert (Ax[0] == D @ x).all() and (Ax[1] == 2.0 * D @ (3.0 * x)).all() and (Ax[2] == -(D @ x)).all()
            assert (ATy[0] == D.T @ y).all() and (ATy[1] == 2.0 * D.T @ (-2.0 * y)).all()
            s.close()
