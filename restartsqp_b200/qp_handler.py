"""Batched QPhandler: the host-side mirror of src/QPhandler.cpp / include/sqphot/QPhandler.hpp.

It maps the current NLP iterates of `batch` independent instances (per-instance trust-region radius
delta and penalty rho, SoA arrays [batch][...]) to QP data, owns one CudaQPInterface backend and
exposes the reference's method names.  The bound/gradient construction of set_bounds / set_g /
update_* runs as device kernels behind sqpb200_qphandler_bounds / sqpb200_qphandler_g.
"""
import numpy as np

from . import _capi as capi
from .qp_interface import CudaQPInterface
from .qore_layout import CudaQOREInterface
from .sqp_types import (IdentityInfo, NLPInfo, Options, QPType, Solver, Stats, SpTripletMat, QP_NOT_OPTIMAL, SQRT_M_EPS,
                        ActiveType)


class QPhandler:
    def __init__(self, nlp_info: NLPInfo, qptype: QPType, options: Options = None, batch=1, device=0, backend=None,
                 refresh_ubA=False, **kw):
        self.nlp_info_ = nlp_info
        self.options = options if options is not None else Options()
        self.batch = batch
        n, m = nlp_info.nVar, nlp_info.nCon
        self.nConstr_QP_ = m              # src/QPhandler.cpp:39
        self.nVar_QP_ = n + 2 * m         # :40
        # I_info_A_: two identity blocks [J I -I], src/QPhandler.cpp:41-51
        self.I_info_A_ = IdentityInfo(irow=np.array([1, 1], np.int32), jcol=np.array([n + 1, n + m + 1], np.int32),
                                      size=np.array([m, m], np.int32), value=np.array([1.0, -1.0]))
        # the backend switch of src/QPhandler.cpp:58-76; `backend` lets a caller plug another object with the
        # CudaQPInterface method set (the tests plug the CPU oracle there to check the host logic without a GPU)
        # QPsolverChoice == QORE selects the QORE-layout member of the family (row-compressed matrices, stacked bounds:
        # CudaQOREInterface), every other choice the qpOASES-layout one; both run on the same CUDA kernels
        choice = self.options.LPsolverChoice if QPType(qptype) == QPType.LP else self.options.QPsolverChoice
        self.QPsolverChoice_ = Solver.QORE if isinstance(backend, CudaQOREInterface) or (backend is None and choice == Solver.QORE) \
            else choice
        if backend is not None:
            self.solverInterface_ = backend
        elif self.QPsolverChoice_ == Solver.QORE:
            self.solverInterface_ = CudaQOREInterface(nlp_info, qptype, self.options, batch=batch, device=device, **kw)
        else:
            self.solverInterface_ = CudaQPInterface(nlp_info, qptype, self.options, batch=batch, device=device, **kw)
        self.qptype = QPType(qptype)
        # False: update_bounds leaves ubA stale exactly like the reference's non-QORE branch;
        # True: both sides are refreshed, as its QORE branch does.  The batched SQP driver needs True to make progress
        # on problems with finite upper constraint bounds.
        self.refresh_ubA = bool(refresh_ubA)
        self.qpOptimalStatus_ = None

    # -- helpers
    def _b(self, a, n):
        """[n] or [batch][n] -> contiguous [batch][n] float64 (numpy) or a CUDA tensor passed through."""
        if capi._is_torch(a):
            return a.contiguous()
        a = np.asarray(a, dtype=np.float64)
        return np.ascontiguousarray(np.broadcast_to(a, (self.batch, n)))

    def _bounds(self, mode, delta, x_l, x_u, x_k, c_l=None, c_u=None, c_k=None):
        n, m = self.nlp_info_.nVar, self.nlp_info_.nCon
        if capi._is_torch(x_k):
            import torch
            d = delta if capi._is_torch(delta) else torch.full((self.batch,), float(delta), dtype=torch.float64, device=x_k.device)
        else:
            d = np.ascontiguousarray(np.broadcast_to(np.asarray(delta, dtype=np.float64), (self.batch,)))
        if mode != 2 and m > 0:
            cs = [self._b(c_l, m), self._b(c_u, m), self._b(c_k, m)]
        else:
            cs = [None, None, None]
        self.solverInterface_.qphandler_bounds(mode, n, m, d, self._b(x_l, n), self._b(x_u, n), self._b(x_k, n), *cs)

    # -- src/QPhandler.cpp:167-261
    def set_bounds(self, delta, x_l, x_u, x_k, c_l, c_u, c_k):
        self._bounds(0, delta, x_l, x_u, x_k, c_l, c_u, c_k)

    # -- src/QPhandler.cpp:342-419 (qpOASES branch: lbA refreshed, ubA not; QORE branch :369-383: both sides)
    def update_bounds(self, delta, x_l, x_u, x_k, c_l, c_u, c_k):
        both = self.refresh_ubA or self.QPsolverChoice_ == Solver.QORE
        self._bounds(3 if both else 1, delta, x_l, x_u, x_k, c_l, c_u, c_k)

    # -- src/QPhandler.cpp:533-567
    def update_delta(self, delta, x_l, x_u, x_k):
        self._bounds(2, delta, x_l, x_u, x_k)

    def _g(self, grad, rho):
        n, m = self.nlp_info_.nVar, self.nlp_info_.nCon
        gr = None if grad is None else self._b(grad, n)
        if rho is None:
            rh = None
        elif capi._is_torch(rho):
            rh = rho.contiguous()
        else:
            rh = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (self.batch,)))
        self.solverInterface_.qphandler_g(n, m, gr, rh)

    # -- src/QPhandler.cpp:272-297 (and the LP overload :657-660 when grad is None)
    def set_g(self, grad, rho=None):
        if rho is None and np.ndim(grad) == 0:
            grad, rho = None, grad
        self._g(grad, rho)

    def update_penalty(self, rho):  # :430-441
        self._g(None, rho)

    def update_grad(self, grad):  # :450-463
        self._g(grad, None)

    def set_H(self, hessian: SpTripletMat):  # :310-318
        self.solverInterface_.set_H(hessian)

    def set_A(self, jacobian: SpTripletMat):  # :326-334
        self.solverInterface_.set_A(jacobian, self.I_info_A_)

    update_H = set_H  # :508-517
    update_A = set_A  # :520-530

    # -- src/QPhandler.cpp:470-499
    def solveQP(self, stats: Stats = None, options: Options = None, active_mask=None):
        self.solverInterface_.optimizeQP(stats, active_mask)
        ok = self.test_optimality()
        if self.batch == 1 and not bool(ok[0]):
            raise QP_NOT_OPTIMAL("KKT error %g > 1e-6" % self.qpOptimalStatus_["KKT_error"][0])
        return ok

    # -- include/sqphot/QPhandler.hpp:66-68
    def solveLP(self, stats: Stats = None, active_mask=None):
        self.solverInterface_.optimizeLP(stats, active_mask)

    def test_optimality(self):  # :580-587
        self.qpOptimalStatus_ = self.solverInterface_.get_optimality_status()
        return self.qpOptimalStatus_["KKT_error"] <= 1.0e-6

    def get_objective(self):  # :502-505
        return self.solverInterface_.get_obj_value()

    def get_optimal_solution(self):  # :128-130
        return self.solverInterface_.get_optimal_solution()

    def get_multipliers_bounds(self):  # :141-143
        return self.solverInterface_.get_multipliers_bounds()

    def get_multipliers_constr(self):  # :146-148
        return self.solverInterface_.get_multipliers_constr()

    def get_status(self):  # :575-577
        return self.solverInterface_.get_status()

    def get_QpOptimalStatus(self):  # :596-598
        return self.qpOptimalStatus_

    def get_infea_measure_model(self):  # :592-594: oneNorm of the slack part of x
        x = self.get_optimal_solution()
        n = self.nlp_info_.nVar
        s = np.zeros(self.batch)
        for i in range(n, self.nVar_QP_):  # index order of Utils oneNorm (src/Utils.cpp:65-72)
            s = s + np.abs(x[:, i])
        return s

    def get_active_set(self, x=None, Ax=None):
        """src/QPhandler.cpp:600-655 (geometric active set, tolerance sqrt_m_eps; reads ubA for both
        constraint sides: quirk 3).  Returns (A_c, A_b)."""
        si = self.solverInterface_
        lb, ub = si.getLb(), si.getUb()
        if x is None:
            x = self.get_optimal_solution()
        if Ax is None:
            Ax = si.spmv(capi.MAT_A, x)
        def classify(v, lo, hi):
            at_lo, at_hi = np.abs(v - lo) < SQRT_M_EPS, np.abs(hi - v) < SQRT_M_EPS
            out = np.full(v.shape, int(ActiveType.INACTIVE), np.int32)
            out[at_hi] = int(ActiveType.ACTIVE_ABOVE)
            out[at_lo] = int(ActiveType.ACTIVE_BELOW)
            out[at_lo & at_hi] = int(ActiveType.ACTIVE_BOTH_SIDE)
            return out
        if self.QPsolverChoice_ == Solver.QORE:  # :626-638: constraint bounds are the tail of the stacked lb / ub
            nV = self.nVar_QP_
            return classify(Ax, lb[:, nV:], ub[:, nV:]), classify(x, lb[:, :nV], ub[:, :nV])
        ubA = si.getUbA()
        return classify(Ax, ubA, ubA), classify(x, lb, ub)

    def WriteQPData(self, filename):  # :569-573
        self.solverInterface_.WriteQPDataToFile(filename)
