"""CudaQOREInterface: the QORE-layout member of the plugin family.

The reference's QOREInterface (include/sqphot/QOREInterface.hpp:30-252, src/QOREInterface.cpp) and its qpOASESInterface
implement the same QPSolverInterface but exchange data in two layouts:

    qpOASES layout                              QORE layout
    A, H column-compressed                      A, H row-compressed (QPSetData, src/QOREInterface.cpp:89-90)
    lb, ub [nV]; lbA, ubA [nC]                  lb, ub [nV + nC] = [variable bounds ; constraint bounds] (:102, :191-219)
    x [nV], y [nV + nC]                         primalsol [nV + nC] = [x ; A x], dualsol [nV + nC] (:120-122)
    working set +1 upper / -1 lower             "workingset" [nV + nC], -1 upper / +1 lower (:440-492)
    set_lbA / set_ubA                           no-ops (include/sqphot/QOREInterface.hpp:180-183)
    test_optimality tolerance 1e-6              1e-5 (:395)

This class offers the QORE layout, with the reference's method names, on the CUDA backend (`batch` instances per object like
CudaQPInterface).  Layout translation happens behind the C ABI (sqpb200_set_structure_*_csr, sqpb200_set_values_csr,
sqpb200_set_bounds_stacked, sqpb200_get_solution_stacked of include/sqpb200.h); nothing numerical runs in this module.
`backend` lets the tests plug a CPU twin with the same method set.
"""
import numpy as np

from . import _capi as capi
from .qp_interface import CudaQPInterface
from .sqp_types import (ActiveType, Exitflag, IdentityInfo, Options, QPType, Stats, SpTripletMat, QP_NOT_OPTIMAL,
                        LP_NOT_OPTIMAL, SQRT_M_EPS, INVALID_WORKING_SET, INF)


class CudaQOREInterface:
    """Batched QP/LP backend in the QORE layout.  nVar_QP_ = nV, nConstr_QP_ = nC (src/QOREInterface.cpp:20-21)."""

    KKT_TOL = 1.0e-5  # src/QOREInterface.cpp:395

    def __init__(self, nlp_info=None, qptype=QPType.QP, options=None, batch=1, device=0, nV=None, nC=None, backend=None, **kw):
        self.options = options if options is not None else Options()
        self.inner = backend if backend is not None else CudaQPInterface(nlp_info, qptype, self.options, batch=batch, device=device,
                                                                          nV=nV, nC=nC, **kw)
        self.nV_, self.nC_, self.batch, self.qptype = self.inner.nV_, self.inner.nC_, self.inner.batch, QPType(qptype)
        self._sol = None  # (primal, dual, workingset) of the last solve, fetched on first use

    def close(self):
        self.inner.close()

    # the C handle and the structure flags of the wrapped backend, for callers that drive the C ABI directly (sqp_device.py)
    h = property(lambda self: self.inner.h)
    _A_set = property(lambda self: self.inner._A_set)
    _H_set = property(lambda self: self.inner._H_set)

    # ------------------------------------------------------------------ setters (include/sqphot/QOREInterface.hpp:142-186)
    def _stacked(self, which, a0, a1):
        n = self.nV_ + self.nC_
        if a1 is None:  # vector form: [nV+nC] (shared) or [batch][nV+nC]
            self.inner.set_bounds_stacked(**{which: a0})
            return
        loc = int(a0)  # (location, value) form: value scalar or [batch]
        if not 0 <= loc < n:
            raise IndexError("location %d outside [0, %d)" % (loc, n))
        a1 = np.maximum(a1, -INF) if which == "lb" else np.minimum(a1, INF)  # include/sqphot/QOREInterface.hpp:147-155
        if loc < self.nV_:
            (self.inner.set_lb if which == "lb" else self.inner.set_ub)(loc, a1)
        else:
            (self.inner.set_lbA if which == "lb" else self.inner.set_ubA)(loc - self.nV_, a1)

    def set_lb(self, a0, a1=None):
        self._stacked("lb", a0, a1)

    def set_ub(self, a0, a1=None):
        self._stacked("ub", a0, a1)

    def set_g(self, a0, a1=None):
        self.inner.set_g(a0, a1 if a1 is None else np.minimum(a1, INF))  # :142-145

    def set_lbA(self, *a):  # include/sqphot/QOREInterface.hpp:180-183: the constraint bounds live in lb / ub
        pass

    set_ubA = set_lbA

    def set_A(self, rhs: SpTripletMat, I_info: IdentityInfo = None):
        self.inner.set_A_csr(rhs, I_info)  # src/QOREInterface.cpp:643-650

    def set_H(self, rhs: SpTripletMat):
        self.inner.set_H_csr(rhs)  # :652-659

    def set_csr(self, which, rowptr, colidx, vals):
        self.inner.set_csr(which, rowptr, colidx, vals)

    def set_csr_values(self, which, vals):
        self.inner.set_csr_values(which, vals)

    def reset_constraints(self):  # include/sqphot/QOREInterface.hpp:185-188
        self.inner.set_bounds_stacked(np.zeros(self.nV_ + self.nC_), np.zeros(self.nV_ + self.nC_))

    # QPhandler's batched data kernels (src/QPhandler.cpp:225-260, 369-383): in the QORE branch update_bounds rewrites both
    # sides of the constraint bounds, so the stale-ubA mode of the qpOASES branch (mode 1) never applies here
    def qphandler_bounds(self, mode, n, m, delta, x_l, x_u, x_k, c_l=None, c_u=None, c_k=None):
        self.inner.qphandler_bounds(3 if mode == 1 else mode, n, m, delta, x_l, x_u, x_k, c_l, c_u, c_k)

    def qphandler_g(self, n, m, grad, rho):
        self.inner.qphandler_g(n, m, grad, rho)

    # ------------------------------------------------------------------ solve (src/QOREInterface.cpp:77-185, 607-629)
    def optimizeQP(self, stats: Stats = None, active_mask=None, maxiter=0):
        """QPSetData + QPOptimize + handle_error (the infeasible branch restarts from the slack-feasible point, :607-629: the
        kernel's recovery path does exactly that).  With batch == 1 raises QP_NOT_OPTIMAL like the reference (:113-115)."""
        self._sol = None
        try:
            self.inner.optimizeQP(stats, active_mask, maxiter)
        except QP_NOT_OPTIMAL:
            raise QP_NOT_OPTIMAL("QP solver reports status %d" % int(self.get_status()[0]))

    def optimizeLP(self, stats: Stats = None, active_mask=None, maxiter=0):
        self._sol = None
        try:
            self.inner.optimizeLP(stats, active_mask, maxiter)
        except LP_NOT_OPTIMAL:
            raise LP_NOT_OPTIMAL("LP solver reports status %d" % int(self.get_status()[0]))

    # ------------------------------------------------------------------ getters
    def _solution(self):
        if self._sol is None:
            self._sol = self.inner.get_solution_stacked()
        return self._sol

    def get_primal_stacked(self):
        """x_qp_ as QORE returns it: [x ; A x] (src/QOREInterface.cpp:120)."""
        return self._solution()[0]

    def get_optimal_solution(self):  # include/sqphot/QOREInterface.hpp:76-78: x_qp_->values(), callers read the first nV
        return self._solution()[0][:, :self.nV_]

    def get_multipliers_bounds(self):  # :83-85
        return self._solution()[1][:, :self.nV_]

    def get_multipliers_constr(self):  # :94-96: y_qp_->values() + nVar_QP_
        return self._solution()[1][:, self.nV_:]

    def get_obj_value(self):
        """0.5 x'Hx + g'x evaluated from the solution (src/QOREInterface.cpp:418-422), not asked of the solver."""
        x = np.ascontiguousarray(self.get_optimal_solution())
        gx = self.inner.vector_times(self.inner.getG(), x)
        if self.qptype == QPType.LP or not getattr(self.inner, "_H_set", True):
            return gx
        return self.inner.vector_times(self.inner.spmv(capi.MAT_H, x), x) * 0.5 + gx  # Hx->times(x_qp_)*0.5 + g_->times(x_qp_)

    def get_status(self):
        """src/QOREInterface.cpp:425-438: QORE knows optimal / iteration limit / infeasible / unbounded; every other state is
        QPERROR_UNKNOWN."""
        st = np.asarray(self.inner.get_status()).copy()
        known = (Exitflag.QP_OPTIMAL, Exitflag.QPERROR_EXCEED_MAX_ITER, Exitflag.QPERROR_INFEASIBLE, Exitflag.QPERROR_UNBOUNDED)
        st[~np.isin(st, [int(k) for k in known])] = int(Exitflag.QPERROR_UNKNOWN)
        return st

    def get_iterations(self):
        return self.inner.get_iterations()

    def get_working_set_raw(self):
        """QORE's "workingset" vector [batch][nV+nC] (src/QOREInterface.cpp:441): -1 upper, +1 lower, 0 inactive."""
        return self._solution()[2]

    def get_working_set(self):
        """(W_constr, W_bounds) as ActiveType values: src/QOREInterface.cpp:440-492 applied to the stacked vectors, the
        misplaced parenthesis of its constraint half (:468, :474: fabs(x - lb < sqrt_m_eps)) included."""
        pr, _, ws = self._solution()
        lb, ub = self.inner.get_bounds_stacked()
        if not np.isin(ws, (-1, 0, 1)).all():
            raise INVALID_WORKING_SET("invalid working set entry")
        nV = self.nV_
        W = np.full(ws.shape, int(ActiveType.INACTIVE), np.int32)
        near_lb = np.empty(ws.shape, bool)
        near_ub = np.empty(ws.shape, bool)
        near_lb[:, :nV] = np.abs(pr[:, :nV] - lb[:, :nV]) < SQRT_M_EPS
        near_ub[:, :nV] = np.abs(pr[:, :nV] - ub[:, :nV]) < SQRT_M_EPS
        near_lb[:, nV:] = pr[:, nV:] - lb[:, nV:] < SQRT_M_EPS  # fabs(bool): true whenever the difference is below the tolerance
        near_ub[:, nV:] = pr[:, nV:] - ub[:, nV:] < SQRT_M_EPS
        up, lo = ws == -1, ws == 1
        W[up] = int(ActiveType.ACTIVE_ABOVE)
        W[up & near_lb] = int(ActiveType.ACTIVE_BOTH_SIDE)
        W[lo] = int(ActiveType.ACTIVE_BELOW)
        W[lo & near_ub] = int(ActiveType.ACTIVE_BOTH_SIDE)
        return W[:, nV:], W[:, :nV]

    def get_optimality_status(self, recompute=False):
        """OptimalityStatus of QOREInterface::test_optimality (src/QOREInterface.cpp:222-409).  Its four sums run over the
        stacked vectors in the order bounds, then constraints, with x_qp_(nV + i) = (A x)_i as the constraint activity: term by
        term the sums of the qpOASES twin (src/qpOASESInterface.cpp:498-684), which the solve kernel's epilogue evaluates."""
        return self.inner.get_optimality_status(recompute)

    def test_optimality(self, recompute=False):
        return self.get_optimality_status(recompute)["KKT_error"] <= self.KKT_TOL

    def getLb(self):
        return self.inner.get_bounds_stacked()[0]

    def getUb(self):
        return self.inner.get_bounds_stacked()[1]

    def getLbA(self):  # include/sqphot/QOREInterface.hpp:124-130 throws: the QORE layout has no separate constraint bounds
        raise AttributeError("the QORE layout has no lbA: constraint bounds are lb[nV:]")

    getUbA = getLbA

    def getG(self):
        return self.inner.getG()

    def getA(self):
        return self.inner.get_csr(capi.MAT_A)

    def getH(self):
        return self.inner.get_csr(capi.MAT_H)

    def spmv(self, which, x, transpose=False):
        return self.inner.spmv(which, x, transpose)

    def launch_count(self):
        return self.inner.launch_count()

    def WriteQPDataToFile(self, filename, instance=0):
        """The `.log` dump of src/QOREInterface.cpp:582-598: sizes, lb, ub, g, A and H row-compressed, one number per line."""
        A = self.getA()
        H = self.getH() if getattr(self.inner, "_H_set", True) and self.qptype != QPType.LP else None
        lb, ub, g = self.getLb(), self.getUb(), self.getG()
        with open(filename, "w") as f:
            for k in (self.nV_, self.nC_, len(A["ColIndex"]), 0 if H is None else len(H["ColIndex"])):
                f.write("%d\n" % k)
            for v in (lb[instance], ub[instance], g[instance]):
                for t in v:
                    f.write("%23.16e\n" % t)
            for M in (A, H):
                if M is None:
                    f.write("0\n" * (self.nV_ + 1))
                    continue
                for t in M["RowIndex"]:
                    f.write("%d\n" % t)
                for t in M["ColIndex"]:
                    f.write("%d\n" % t)
                for t in M["MatVal"][instance]:
                    f.write("%23.16e\n" % t)


def read_qore_log_raw(path):
    """The `.log` layout as it is, without the row- to column-compressed conversion of the replay driver: what
    test/QPsolvers_testers.cpp:48-150 reads and hands to the QORE data constructor (:74-75, :172-175)."""
    it = iter(open(path).read().split())
    nV, nC, zA, zH = (int(next(it)) for _ in range(4))
    f = lambda n: np.array([float(next(it)) for _ in range(n)], dtype=np.float64)
    i = lambda n: np.array([int(next(it)) for _ in range(n)], dtype=np.int32)
    lb, ub, g = f(nV + nC), f(nV + nC), f(nV)
    A_rp, A_ci, A_v = i(nC + 1), i(zA), f(zA)
    H_rp, H_ci, H_v = i(nV + 1), i(zH), f(zH)
    return dict(nV=nV, nC=nC, lb=lb, ub=ub, g=g, A_rowptr=A_rp, A_colidx=A_ci, A_val=A_v, H_rowptr=H_rp, H_colidx=H_ci, H_val=H_v)


def replay_qore(q, batch=1, device=0, options=None, backend=None, **kw):
    """QOREInterface(H, A, g, lb, ub, options) (src/QOREInterface.cpp:36-60) for a dumped QP in the `.log` layout (a path or the
    dict of read_qore_log_raw), every instance of the batch holding the same data."""
    if isinstance(q, str):
        q = read_qore_log_raw(q)
    s = CudaQOREInterface(nV=q["nV"], nC=q["nC"], qptype=QPType.QP, options=options, batch=batch, device=device, backend=backend, **kw)
    s.set_csr(capi.MAT_A, q["A_rowptr"], q["A_colidx"], np.asarray(q["A_val"], dtype=np.float64))
    s.set_csr(capi.MAT_H, q["H_rowptr"], q["H_colidx"], np.asarray(q["H_val"], dtype=np.float64))
    s.set_g(np.asarray(q["g"], dtype=np.float64))
    s.set_lb(np.asarray(q["lb"], dtype=np.float64))
    s.set_ub(np.asarray(q["ub"], dtype=np.float64))
    return s
