"""The QORE-layout member of the plugin family (restartsqp_b200/qore_layout.py, mirror of include/sqphot/QOREInterface.hpp and
src/QOREInterface.cpp) on CPU: the host class is driven through the oracle-backed twin of the
backend (tests/oracle_backend.py), so what is tested here is the layout logic -- stacked bounds, [x ; A x], QORE's working-set
sign and its translation, status mapping, tolerances, the QORE branches of QPhandler -- not the arithmetic.  The GPU twin of
this file is tests/test_gpu_qore.py."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import restartsqp_b200 as r
from restartsqp_b200 import capi, qp_dump
from restartsqp_b200.qore_layout import CudaQOREInterface, read_qore_log_raw, replay_qore
from restartsqp_b200.sqp_driver import BatchedSQP, HS071
from restartsqp_b200.sqp_types import SQRT_M_EPS
from oracle_backend import OracleQPInterface
from helpers import load_l0_golden, load_qore_golden, load_qp_fixtures, load_qore_raw_fixtures, random_l1_qp, is_symmetric_fixture, oracle_solve

L0 = {c["name"]: c for c in load_l0_golden()}


def qore_on_oracle(nV, nC, batch=1, qptype=r.QPType.QP):
    return CudaQOREInterface(nV=nV, nC=nC, qptype=qptype, batch=batch, backend=OracleQPInterface(nV=nV, nC=nC, qptype=int(qptype), batch=batch))


def reference_translation(ws, primal, lb, ub, nV):
    """src/QOREInterface.cpp:440-492, entry by entry."""
    W = np.zeros(len(ws), np.int32)
    for i in range(len(ws)):
        if i < nV:
            near_lb, near_ub = abs(primal[i] - lb[i]) < SQRT_M_EPS, abs(primal[i] - ub[i]) < SQRT_M_EPS
        else:  # fabs(x - lb < sqrt_m_eps): the comparison is inside the fabs
            near_lb, near_ub = abs(float(primal[i] - lb[i] < SQRT_M_EPS)) != 0, abs(float(primal[i] - ub[i] < SQRT_M_EPS)) != 0
        if ws[i] == -1:
            W[i] = -99 if near_lb else 1
        elif ws[i] == 1:
            W[i] = -99 if near_ub else -1
    return W


@pytest.mark.parametrize("q", load_qore_golden()[:6], ids=lambda q: q["name"])
def test_triplets_become_the_reference_row_compressed_arrays(q):
    c = L0[q["name"]]
    n, m = c["n"], c["m"]
    nV, nC = n + 2 * m, m
    s = qore_on_oracle(nV, nC, batch=2)
    info = r.IdentityInfo(irow=np.array([1, 1], np.int32), jcol=np.array([n + 1, n + m + 1], np.int32),
                          size=np.array([m, m], np.int32), value=np.array([1.0, -1.0]))
    s.set_A(r.SpTripletMat(c["J_row1"], c["J_col1"], np.array(c["J_val"]), nC, nV, False), info)
    s.set_H(r.SpTripletMat(c["H_row1"], c["H_col1"], np.array(c["H_val"]), nV, nV, True))
    A, H = s.getA(), s.getH()
    assert A["RowIndex"].tolist() == q["A_rowptr"] and A["ColIndex"].tolist() == q["A_colidx"] and A["order"].tolist() == q["A_order"]
    assert A["MatVal"][1].tolist() == q["A_val"]
    assert H["RowIndex"].tolist() == q["H_rowptr"] and H["ColIndex"].tolist() == q["H_colidx"] and H["order"].tolist() == q["H_order"]
    assert H["MatVal"][0].tolist() == q["H_val"]
    s.set_A(r.SpTripletMat(c["J_row1"], c["J_col1"], np.array(c["J_val2"]), nC, nV, False), info)  # later call: value refresh
    s.set_H(r.SpTripletMat(c["H_row1"], c["H_col1"], np.array(c["H_val2"]), nV, nV, True))
    assert s.getA()["MatVal"][0].tolist() == q["A_val2"] and s.getH()["MatVal"][1].tolist() == q["H_val2"]


def load_both_layouts(p, batch, g):
    """One random l1-penalty QP in both layouts on the oracle twin."""
    nV, nC = p["nV"], p["nC"]
    Acsr, Hcsr = sp.csr_matrix(p["A"]), sp.csr_matrix(p["H"])
    Acsr.sort_indices(); Hcsr.sort_indices()
    Acsc, Hcsc = sp.csc_matrix(p["A"]), sp.csc_matrix(p["H"])
    Acsc.sort_indices(); Hcsc.sort_indices()
    q = qore_on_oracle(nV, nC, batch)
    q.set_csr(capi.MAT_A, Acsr.indptr, Acsr.indices, Acsr.data)
    q.set_csr(capi.MAT_H, Hcsr.indptr, Hcsr.indices, Hcsr.data)
    q.set_g(g)
    q.set_lb(np.concatenate([p["lb"], p["lbA"]]))
    q.set_ub(np.concatenate([p["ub"], p["ubA"]]))
    o = OracleQPInterface(nV=nV, nC=nC, batch=batch)
    o.A = dict(p=Acsc.indptr.astype(np.int32), i=Acsc.indices.astype(np.int32), zJ=Acsc.nnz)
    o.H = dict(p=Hcsc.indptr.astype(np.int32), i=Hcsc.indices.astype(np.int32))
    o.Av, o.Hv = np.tile(Acsc.data, (batch, 1)), np.tile(Hcsc.data, (batch, 1))
    o.set_g(g); o.set_lb(p["lb"]); o.set_ub(p["ub"]); o.set_lbA(p["lbA"]); o.set_ubA(p["ubA"])
    return q, o


@pytest.mark.parametrize("seed", range(6))
def test_qore_layout_returns_what_the_qpoases_layout_returns(seed):
    rng = np.random.default_rng(400 + seed)
    n, m, B = int(rng.integers(2, 7)), int(rng.integers(0, 5)), 3
    p = random_l1_qp(rng, n, m)
    nV, nC = p["nV"], p["nC"]
    g = np.tile(p["g"], (B, 1))
    g[:, :n] += 0.3 * rng.standard_normal((B, n))
    q, o = load_both_layouts(p, B, g)
    q.optimizeQP(); o.optimizeQP()
    assert (q.get_status() == 20).all() and (o.get_status() == 20).all()
    pr = q.get_primal_stacked()
    assert pr.shape == (B, nV + nC)
    assert (pr[:, :nV] == o.get_optimal_solution()).all() and (q.get_optimal_solution() == o.get_optimal_solution()).all()
    assert np.allclose(pr[:, nV:], o.get_optimal_solution() @ p["A"].T, rtol=0, atol=1e-12)   # the tail is the constraint activity A x
    assert (q.get_multipliers_bounds() == o.get_multipliers_bounds()).all()
    assert (q.get_multipliers_constr() == o.get_multipliers_constr()).all()
    raw = q.get_working_set_raw()
    assert (raw == -np.hstack([o.wb, o.wc])).all()  # QORE: -1 upper, +1 lower
    lb, ub = q.getLb(), q.getUb()
    assert lb.shape == (B, nV + nC) and (lb[:, nV:] == p["lbA"]).all() and (ub[:, :nV] == p["ub"]).all()
    Wc, Wb = q.get_working_set()
    for b in range(B):
        W = reference_translation(raw[b], pr[b], lb[b], ub[b], nV)
        assert (Wb[b] == W[:nV]).all() and (Wc[b] == W[nV:]).all()
    assert np.abs(q.get_obj_value() - o.get_obj_value()).max() < 1e-9 * max(1.0, np.abs(o.get_obj_value()).max())
    assert q.test_optimality().all() and q.KKT_TOL == 1.0e-5
    assert (q.get_iterations() == o.get_iterations()).all()
    with pytest.raises(AttributeError):
        q.getLbA()


def test_location_value_setters_address_the_stacked_vectors():
    """set_lb(location, value): locations >= nV are constraint bounds (src/QPhandler.cpp:230-233, 377-382); set_lbA / set_ubA
    are no-ops (include/sqphot/QOREInterface.hpp:180-183); reset_constraints zeroes both stacked vectors."""
    q = qore_on_oracle(5, 2, batch=3)
    q.set_lb(np.arange(7.0)); q.set_ub(np.arange(7.0) + 10)
    q.set_lb(1, -4.0); q.set_ub(6, np.array([1.0, 2.0, 3.0])); q.set_lb(5, 0.5)
    q.set_lbA(0, 99.0); q.set_ubA(np.ones(2))
    lb, ub = q.getLb(), q.getUb()
    assert lb[0].tolist() == [0, -4, 2, 3, 4, 0.5, 6] and ub[:, 6].tolist() == [1, 2, 3] and ub[1, :6].tolist() == [10, 11, 12, 13, 14, 15]
    with pytest.raises(IndexError):
        q.set_lb(7, 0.0)
    q.reset_constraints()
    assert not q.getLb().any() and not q.getUb().any()


def test_status_mapping_of_the_qore_backend():
    """src/QOREInterface.cpp:425-438: anything but optimal / iteration limit / infeasible / unbounded reads QPERROR_UNKNOWN."""
    q = qore_on_oracle(3, 1, batch=6)
    q.inner.status[:] = [20, 21, 22, 23, 24, 25]
    assert q.get_status().tolist() == [20, 30, 22, 23, 24, 30]


def test_dump_round_trip_in_the_log_layout(tmp_path):
    """WriteQPDataToFile writes the `.log` layout of src/QOREInterface.cpp:582-598; read back without conversion and replayed
    through the QORE data constructor it gives the same solve; the file equals what qp_dump writes for the same QP."""
    fx = [q for q in load_qp_fixtures() if q["name"] == "QORE_hs064"][0]
    p1 = str(tmp_path / "a.log")
    qp_dump.write_qore_log(p1, fx)
    raw = read_qore_log_raw(p1)
    assert raw["nV"] == fx["nV"] and raw["lb"].tolist() == list(fx["lb"]) + list(fx["lbA"])
    s = replay_qore(p1, batch=2, backend=OracleQPInterface(nV=fx["nV"], nC=fx["nC"], batch=2))
    s.optimizeQP()
    p2 = str(tmp_path / "b.log")
    s.WriteQPDataToFile(p2, instance=1)
    assert open(p1).read() == open(p2).read()
    from oracle import oracle_py as orc
    from helpers import oracle_solve
    o = oracle_solve(orc, dict(nV=fx["nV"], nC=fx["nC"], g=np.array(fx["g"]), lb=np.array(fx["lb"]), ub=np.array(fx["ub"]),
                               lbA=np.array(fx["lbA"]), ubA=np.array(fx["ubA"])),
                     Acsc=(np.array(fx["A_colptr"], np.int32), np.array(fx["A_rowidx"], np.int32), np.array(fx["A_val"])),
                     Hcsc=(np.array(fx["H_colptr"], np.int32), np.array(fx["H_rowidx"], np.int32), np.array(fx["H_val"])))
    assert int(s.get_status()[0]) == o["status"] and (s.get_optimal_solution()[0] == o["x"]).all()
    assert (s.get_working_set_raw()[1] == -np.concatenate([o["wb"], o["wc"]])).all()


def qore_handler_factory(batch, options):
    def make(info, qptype):
        be = CudaQOREInterface(info, qptype, options, batch=batch, backend=OracleQPInterface(info, qptype, options, batch=batch))
        return r.QPhandler(info, qptype, options, batch=batch, backend=be)
    return make


def test_qphandler_qore_branches():
    """src/QPhandler.cpp:225-260 / 369-383: the QORE branch writes the constraint bounds into the tail of lb / ub and refreshes
    BOTH sides in update_bounds (the qpOASES branch leaves ubA stale); get_active_set reads the stacked bounds (:626-638)."""
    info = r.NLPInfo(nCon=2, nVar=3)
    B = 2
    h = qore_handler_factory(B, r.Options())(info, r.QPType.QP)
    assert h.QPsolverChoice_ == r.Solver.QORE
    x_l, x_u, x_k = np.full(3, -5.0), np.full(3, 5.0), np.array([0.5, -1.0, 4.8])
    c_l, c_u, c_k = np.array([0.0, -1e18]), np.array([2.0, 1.0]), np.array([0.25, 0.5])
    h.set_bounds(np.array([1.0, 0.1]), x_l, x_u, x_k, c_l, c_u, c_k)
    lb, ub = h.solverInterface_.getLb(), h.solverInterface_.getUb()
    nV = 3 + 4
    assert lb.shape == (B, nV + 2)
    assert lb[0, :3].tolist() == [-1.0, -1.0, -1.0] and ub[0, :3].tolist() == [1.0, 1.0, 5.0 - 4.8]
    assert ub[1, :3].tolist() == [0.1, 0.1, 0.1] and (lb[:, 3:nV] == 0).all() and (ub[:, 3:nV] == 1e18).all()
    assert lb[0, nV:].tolist() == [-0.25, -1e18 - 0.5] and ub[0, nV:].tolist() == [1.75, 0.5]
    h.update_bounds(np.array([1.0, 0.1]), x_l, x_u, x_k, c_l, c_u, c_k + 1.0)
    lb, ub = h.solverInterface_.getLb(), h.solverInterface_.getUb()
    assert lb[1, nV:].tolist() == [-1.25, -1e18 - 1.5] and ub[1, nV:].tolist() == [0.75, -0.5]   # upper side refreshed too
    # geometric active set with stacked bounds
    x = np.tile(np.concatenate([[1.0, 0.0, 0.2], np.zeros(4)]), (B, 1))
    Ax = np.tile([0.75, 0.0], (B, 1))
    A_c, A_b = h.get_active_set(x, Ax)
    assert A_b[0, :3].tolist() == [1, 0, -99 if abs(0.2 - (5.0 - 4.8)) < SQRT_M_EPS and abs(0.2 + 1.0) < SQRT_M_EPS else 1]
    assert A_c[0].tolist() == [1, 0] and A_c[1].tolist() == [1, 0]


def test_sqp_loop_through_the_qore_layout_equals_the_qpoases_layout():
    """Algorithm::Optimize with QPsolverChoice = QORE (src/QPhandler.cpp:63-64): same iterates as through the qpOASES layout
    with both constraint sides refreshed -- the layouts differ, the QPs do not."""
    rng = np.random.default_rng(3)
    x0 = np.array([1.0, 5.0, 5.0, 1.0])
    B = 4
    starts = np.clip(x0 * (1 + 0.1 * rng.standard_normal((B, 4))), 1.0, 5.0)
    opt = r.Options()
    res_q = BatchedSQP(HS071(), x0=starts, options=opt, make_handler=qore_handler_factory(B, opt)).Optimize()

    def plain(info, qptype):
        return r.QPhandler(info, qptype, opt, batch=B, backend=OracleQPInterface(info, qptype, opt, batch=B), refresh_ubA=True)
    res_p = BatchedSQP(HS071(), x0=starts, options=opt, make_handler=plain).Optimize()
    assert (res_q.exitflag == res_p.exitflag).all() and (res_q.exitflag == 0).all()
    assert (res_q.iters == res_p.iters).all() and (res_q.qp_iter == res_p.qp_iter).all()
    assert (res_q.x == res_p.x).all()


RAW = load_qore_raw_fixtures()


@pytest.mark.parametrize("q", RAW, ids=[q["name"] for q in RAW])
def test_replay_driver_both_arms_agree(q):
    """test/QPsolvers_testers.cpp solves each dumped QP twice: as it is with the QORE-layout backend (:172-200) and, converted
    through the dense matrix (entries with |v| <= 1e-16 dropped, :206-218), with the qpOASES-layout backend (:220-229), and
    prints the two side by side.  Same two arms on the oracle twin: identical status, iteration count and working set; the
    solutions agree to the last bits (the unconverted matrices keep explicit zeros and, in hs104, entries of 1e-17)."""
    from oracle import oracle_py as orc
    f = [c for c in load_qp_fixtures() if c["name"] == q["name"]][0]
    s = replay_qore(q, batch=1, backend=OracleQPInterface(nV=q["nV"], nC=q["nC"], batch=1))
    s.inner.optimizeQP()  # no exception for the reference's own failure case (hs107)
    o = oracle_solve(orc, dict(nV=f["nV"], nC=f["nC"], g=np.array(f["g"]), lb=np.array(f["lb"]), ub=np.array(f["ub"]),
                               lbA=np.array(f["lbA"]), ubA=np.array(f["ubA"])),
                     Acsc=(np.array(f["A_colptr"], np.int32), np.array(f["A_rowidx"], np.int32), np.array(f["A_val"])),
                     Hcsc=(np.array(f["H_colptr"], np.int32), np.array(f["H_rowidx"], np.int32), np.array(f["H_val"])))
    assert int(s.inner.get_status()[0]) == o["status"] and int(s.get_iterations()[0]) == o["iters"]
    assert (s.get_working_set_raw()[0] == -np.concatenate([o["wb"], o["wc"]])).all()
    x = s.get_optimal_solution()[0]
    assert np.abs(x - o["x"]).max() <= 1e-12 * max(1.0, np.abs(o["x"]).max())
