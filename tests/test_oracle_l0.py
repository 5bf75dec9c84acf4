"""Pin the CPU oracle's L0 restatement (oracle/oracle_l0.c) against
  (a) committed golden vectors produced by the reference's own code (tests/golden/l0_golden.json), and
  (b) when oracle/_ref is built (dev container), the reference library itself on fresh random inputs.
Integer arrays and FP64 results must be bit-identical."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from helpers import load_l0_golden

CASES = load_l0_golden()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_csc_assembly_matches_reference_golden(case):
    n, m = case["n"], case["m"]
    nV, nC = n + 2 * m, m
    Ap, Ai, Av, Ao = orc.assemble_A(nC, nV, case["J_row1"], case["J_col1"], case["J_val"], orc.identity_info(n, m))
    assert Ap.tolist() == case["A_colptr"]
    assert Ai.tolist() == case["A_rowidx"]
    assert Ao.tolist() == case["A_order"]
    assert Av.tolist() == case["A_val"]
    Hp, Hi, Hv, Ho = orc.assemble_H(nV, case["H_row1"], case["H_col1"], case["H_val"], True)
    assert Hp.tolist() == case["H_colptr"]
    assert Hi.tolist() == case["H_rowidx"]
    assert Ho.tolist() == case["H_order"]
    assert Hv.tolist() == case["H_cscval"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_value_refresh_matches_reference_golden(case):
    A2 = orc.setmatval_A(case["A_order"], case["J_val2"], case["A_val"])
    assert A2.tolist() == case["A_val2"]
    H2 = orc.setmatval_H(case["H_row1"], case["H_col1"], case["H_order"], case["H_val2"], case["H_cscval"], True)
    assert H2.tolist() == case["H_cscval2"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_spmv_matches_reference_golden(case):
    n, m = case["n"], case["m"]
    nV, nC = n + 2 * m, m
    Ax = orc.csc_times(nC, nV, case["A_colptr"], case["A_rowidx"], case["A_val"], case["x"])
    ATy = orc.csc_times(nC, nV, case["A_colptr"], case["A_rowidx"], case["A_val"], case["yc"], transpose=True)
    Hx = orc.csc_times(nV, nV, case["H_colptr"], case["H_rowidx"], case["H_cscval"], case["x"])
    assert Ax.tolist() == case["Ax"]
    assert ATy.tolist() == case["ATy"]
    assert Hx.tolist() == case["Hx"]
    x = np.array(case["x"])
    assert orc.lib().orc_one_norm(orc._dp(x), nV) == case["one_norm_x"]
    assert orc.lib().orc_inf_norm(orc._dp(x), nV) == case["inf_norm_x"]


def test_hs071_probe_of_survey():
    """The HS071-shaped probe of SURVEY.md section 8c, row-major triplets."""
    c = [c for c in CASES if c["name"] == "hs071_rowmajor"][0]
    assert c["A_colptr"] == [0, 2, 4, 6, 8, 9, 10, 11, 12]
    assert c["A_rowidx"] == [0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1]
    assert c["A_order"] == [0, 2, 4, 6, 1, 3, 5, 7, 8, 9, 10, 11]
    assert c["H_colptr"] == [0, 4, 8, 12, 16, 16, 16, 16, 16]


def test_empty_and_ragged_inputs():
    # no constraints at all (hs038-like): A is 0 x n with no entries
    Ap, Ai, Av, Ao = orc.assemble_A(0, 3, [], [], [], orc.identity_info(3, 0))
    assert Ap.tolist() == [0, 0, 0, 0] and len(Ai) == 0
    # empty columns in the middle and at the end
    Hp, Hi, Hv, Ho = orc.assemble_H(5, [4, 1], [4, 1], [2.0, 3.0], True)
    assert Hp.tolist() == [0, 1, 1, 1, 2, 2] and Hi.tolist() == [0, 3] and Ho.tolist() == [1, 0]
    # duplicate keys: index arrays are well defined, order is the stable one
    colptr, rowidx, val, order = orc.csc_from_entries(2, [1, 1, 1], [1, 1, 2], [1.0, 2.0, 3.0])
    assert colptr.tolist() == [0, 2, 3] and rowidx.tolist() == [0, 0, 0] and order.tolist() == [0, 1, 2]


@pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_against_reference_library_random():
    """Fresh random matrices through the reference's own SpHbMat code (oracle/_ref)."""
    import ctypes as C
    R = orc.ref_lib()
    rng = np.random.default_rng(7)
    ip, dp = orc._ip, orc._dp
    for trial in range(40):
        n, m = int(rng.integers(1, 15)), int(rng.integers(0, 12))
        nV, nC = n + 2 * m, m
        mask = rng.random((m, n)) < 0.5
        rr, cc = np.nonzero(mask)
        perm = rng.permutation(len(rr))
        jr, jc = (rr[perm] + 1).astype(np.int32), (cc[perm] + 1).astype(np.int32)
        jv = rng.integers(1, 10, len(jr)).astype(np.float64)
        iinfo = orc.identity_info(n, m)
        irow, jcol, size, ival = [a.copy() for a in iinfo]
        z = len(jr) + 2 * m
        colptr, rowidx, order, val = np.zeros(nV + 1, np.int32), np.zeros(z, np.int32), np.zeros(z, np.int32), np.zeros(z)
        R.ref_assemble_A(nC, nV, len(jr), ip(jr), ip(jc), dp(jv), 2, ip(irow), ip(jcol), ip(size), dp(ival), None,
                         ip(colptr), ip(rowidx), dp(val), ip(order))
        Ap, Ai, Av, Ao = orc.assemble_A(nC, nV, jr, jc, jv, iinfo)
        assert (Ap == colptr).all() and (Ai == rowidx).all() and (Ao == order).all() and (Av == val).all()
        x, yc = rng.standard_normal(nV), rng.standard_normal(nC)
        y_ref, yt_ref = np.zeros(nC), np.zeros(nV)
        R.ref_csc_times(nC, nV, z, ip(colptr), ip(rowidx), dp(val), dp(x), dp(y_ref))
        R.ref_csc_transposed_times(nC, nV, z, ip(colptr), ip(rowidx), dp(val), dp(yc), dp(yt_ref))
        assert (orc.csc_times(nC, nV, Ap, Ai, Av, x) == y_ref).all()
        assert (orc.csc_times(nC, nV, Ap, Ai, Av, yc, transpose=True) == yt_ref).all()
        assert R.ref_vector_one_norm(dp(x), nV) == orc.lib().orc_one_norm(dp(x), nV)
    assert R.ref_const_INF() == 1.0e18 and R.ref_const_sqrt_m_eps() == 1.0e-8


def test_qphandler_data_construction():
    """src/QPhandler.cpp:185-201, 358-367, 559-564 and quirk 2 (update_bounds leaves ubA stale)."""
    n, m = 3, 2
    x_l, x_u, x_k = np.array([-5.0, 0.0, 1.0]), np.array([5.0, 0.4, 1e18]), np.array([0.0, 0.2, 3.0])
    c_l, c_u, c_k = np.array([1.0, -1e18]), np.array([1.0, 4.0]), np.array([0.5, 1.0])
    lb, ub, lbA, ubA = np.zeros(n + 2 * m), np.zeros(n + 2 * m), np.zeros(m), np.zeros(m)
    orc.qp_bounds(0, n, m, 1.0, x_l, x_u, x_k, c_l, c_u, c_k, lb, ub, lbA, ubA)
    assert lb.tolist() == [-1.0, -0.2, -1.0, 0, 0, 0, 0]
    assert ub.tolist() == [1.0, 0.2, 1.0, 1e18, 1e18, 1e18, 1e18]
    assert lbA.tolist() == [0.5, -1e18 - 1.0] and ubA.tolist() == [0.5, 3.0]
    orc.qp_bounds(1, n, m, 0.5, x_l, x_u, x_k, c_l, c_u, c_k + 1.0, lb, ub, lbA, ubA)
    assert lbA.tolist() == [-0.5, -1e18 - 2.0] and ubA.tolist() == [0.5, 3.0]  # ubA is stale
    assert lb[:3].tolist() == [-0.5, -0.2, -0.5]


def test_working_set_translation_quirk():
    """src/qpOASESInterface.cpp:874/880: a lower-active constraint is always reported ACTIVE_BOTH_SIDE."""
    Wb, Wc = orc.translate_working_set([1, -1, 0], [-1, 1, 0], [1.0, 0.0, 0.5], [2.0, 5.0, 0.0],
                                       [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [2.0, 1.0, -1.0], [3.0, 5.0, 1.0])
    assert Wb.tolist() == [1, -1, 0]
    assert Wc.tolist() == [-99, 1, 0]
