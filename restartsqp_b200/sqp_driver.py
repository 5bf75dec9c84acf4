"""Batched Sl1QP trust-region SQP driver: the host-side mirror of src/Algorithm.cpp.

It holds `batch` independent NLP instances of one model (same sparsity structure, different starting points) with
per-instance trust-region radius, penalty parameter, eps1, merit values, flags and exit codes in SoA arrays, and runs
the reference's outer loop with masked execution: every QP / LP subproblem of the batch is one call into the batched
QPhandler (one kernel launch), instances that have terminated are masked out.

The control flow and the arithmetic follow Algorithm::Optimize and its helpers line by line; the citations name the
reference lines.  The trust-region radius update, the penalty update and the restart logic are unchanged: the CUDA
backend is a drop-in behind QPhandler.

Batch semantics of the backend's init/hotstart state machine: the `Update_A/Update_H` flags of
src/qpOASESInterface.cpp:361-496 are raised per handle, so one outer iteration uses a matrix-changing hot start for the
whole batch as soon as one instance accepted its step (for an instance whose matrices did not change this is the same QP
solved from the same working set with freshly built factors).  The device-resident driver (sqp_device.DeviceBatchedSQP) keeps
that state machine per instance instead.
"""
from dataclasses import dataclass

import numpy as np

from .qp_handler import QPhandler
from .sqp_types import Exitflag, NLPInfo, Options, QPType, SpTripletMat, Stats, INF, QP_NOT_OPTIMAL, LP_NOT_OPTIMAL

# ConstraintType, include/sqphot/Types.hpp:75-81
BOUNDED, EQUAL, BOUNDED_ABOVE, BOUNDED_BELOW, UNBOUNDED = 5, -5, 9, 1, 0


def classify_single_constraint(lo, hi):
    """src/Utils.cpp:29-45, vectorised (including its `upper_bound > INF` test for BOUNDED_BELOW)."""
    out = np.full(lo.shape, UNBOUNDED, dtype=np.int32)
    both = (lo > -INF) & (hi < INF)
    out[both & ((hi - lo) < 1.0e-8)] = EQUAL
    out[both & ~((hi - lo) < 1.0e-8)] = BOUNDED
    out[~both & (lo > -INF) & (hi > INF)] = BOUNDED_BELOW
    out[~both & ~((lo > -INF) & (hi > INF)) & (hi < INF) & (lo < -INF)] = BOUNDED_ABOVE
    return out


class HS071:
    """Hock-Schittkowski 71 (test/CUTE_examples/hs071.nl: n=4, m=2, x0=(1,5,5,1), 1<=x<=5, c1>=25, c2=40), batched.
    Triplets in the order AmplTNLP emits them (Jacobian column-major, Hessian upper triangle by columns), 1-based."""
    n, m = 4, 2
    J_row1 = np.array([1, 2, 1, 2, 1, 2, 1, 2], np.int32)
    J_col1 = np.array([1, 1, 2, 2, 3, 3, 4, 4], np.int32)
    H_row1 = np.array([1, 1, 2, 1, 2, 3, 1, 2, 3, 4], np.int32)
    H_col1 = np.array([1, 2, 2, 3, 3, 3, 4, 4, 4, 4], np.int32)

    def Get_nlp_info(self):
        return NLPInfo(nCon=2, nVar=4, nnz_jac_g=8, nnz_h_lag=10)

    def Get_bounds_info(self):
        return np.ones(4), 5.0 * np.ones(4), np.array([25.0, 40.0]), np.array([1.0e19, 40.0])

    def Get_starting_point(self):
        return np.array([1.0, 5.0, 5.0, 1.0]), np.zeros(2)

    def Eval_f(self, x):
        return x[:, 0] * x[:, 3] * (x[:, 0] + x[:, 1] + x[:, 2]) + x[:, 2]

    def Eval_gradient(self, x):
        g = np.empty_like(x)
        g[:, 0] = x[:, 3] * (2 * x[:, 0] + x[:, 1] + x[:, 2])
        g[:, 1] = x[:, 0] * x[:, 3]
        g[:, 2] = x[:, 0] * x[:, 3] + 1.0
        g[:, 3] = x[:, 0] * (x[:, 0] + x[:, 1] + x[:, 2])
        return g

    def Eval_constraints(self, x):
        return np.stack([x[:, 0] * x[:, 1] * x[:, 2] * x[:, 3], (x * x).sum(1)], 1)

    def Eval_Jacobian(self, x):
        v = np.empty((x.shape[0], 8))
        v[:, 0] = x[:, 1] * x[:, 2] * x[:, 3]; v[:, 1] = 2 * x[:, 0]
        v[:, 2] = x[:, 0] * x[:, 2] * x[:, 3]; v[:, 3] = 2 * x[:, 1]
        v[:, 4] = x[:, 0] * x[:, 1] * x[:, 3]; v[:, 5] = 2 * x[:, 2]
        v[:, 6] = x[:, 0] * x[:, 1] * x[:, 2]; v[:, 7] = 2 * x[:, 3]
        return v

    def Eval_Hessian(self, x, lam):
        """Hessian of f + sum_i lam_i c_i (Ipopt eval_h with obj_factor 1); SQPTNLP passes -multiplier
        (src/SQPTNLP.cpp:124-126), which the driver does before calling this."""
        l1, l2 = lam[:, 0], lam[:, 1]
        v = np.empty((x.shape[0], 10))
        v[:, 0] = 2 * x[:, 3] + 2 * l2                                   # (1,1)
        v[:, 1] = x[:, 3] + l1 * x[:, 2] * x[:, 3]                       # (1,2)
        v[:, 2] = 2 * l2                                                 # (2,2)
        v[:, 3] = x[:, 3] + l1 * x[:, 1] * x[:, 3]                       # (1,3)
        v[:, 4] = l1 * x[:, 0] * x[:, 3]                                 # (2,3)
        v[:, 5] = 2 * l2                                                 # (3,3)
        v[:, 6] = 2 * x[:, 0] + x[:, 1] + x[:, 2] + l1 * x[:, 1] * x[:, 2]  # (1,4)
        v[:, 7] = x[:, 0] + l1 * x[:, 0] * x[:, 2]                       # (2,4)
        v[:, 8] = x[:, 0] + l1 * x[:, 0] * x[:, 1]                       # (3,4)
        v[:, 9] = 2 * l2                                                 # (4,4)
        return v


@dataclass
class SQPResult:
    x: np.ndarray
    obj: np.ndarray
    exitflag: np.ndarray
    iters: np.ndarray
    qp_iter: np.ndarray
    KKT_error: np.ndarray
    rho: np.ndarray
    delta: np.ndarray


class BatchedSQP:
    """Algorithm (include/sqphot/Algorithm.hpp) for a batch of instances.  `make_handler(nlp_info, qptype)` builds the
    QP and LP handlers (default: the CUDA-backed QPhandler)."""

    def __init__(self, nlp, x0=None, options: Options = None, make_handler=None, device=0, dump_dir=None, dump_max=4):
        """dump_dir: where the QP of an instance whose subproblem failed is written in the QORE `.log` layout (the
        `myQP_->WriteQPData(problem_name + "qpdata.log")` of src/Algorithm.cpp:69), at most dump_max files per run."""
        self.dump_dir_, self.dump_left_ = dump_dir, dump_max
        self.nlp_ = nlp
        self.options_ = options if options is not None else Options()
        self.info = nlp.Get_nlp_info()
        n, m = self.info.nVar, self.info.nCon
        self.nVar_, self.nCon_ = n, m
        x_start, lam_start = nlp.Get_starting_point()
        x0 = np.atleast_2d(np.asarray(x_start if x0 is None else x0, dtype=np.float64))
        B = self.batch = x0.shape[0]
        if make_handler is None:
            make_handler = lambda info, qptype: QPhandler(info, qptype, self.options_, batch=B, device=device, refresh_ubA=True)
        self.myQP_ = make_handler(self.info, QPType.QP)  # src/Algorithm.cpp:561-562
        self.myLP_ = make_handler(self.info, QPType.LP)
        self.stats_ = Stats()
        o = self.options_
        # ---- initialization(), src/Algorithm.cpp:438-472
        self.delta_ = np.full(B, float(o.delta))
        self.rho_ = np.full(B, float(o.rho))
        self.eps1_ = np.full(B, float(o.eps1))  # mutated per instance at run time (:984-985)
        xl, xu, cl, cu = nlp.Get_bounds_info()
        self.x_l_, self.x_u_ = np.tile(xl, (B, 1)), np.tile(xu, (B, 1))
        self.c_l_, self.c_u_ = np.tile(cl, (B, 1)).reshape(B, m), np.tile(cu, (B, 1)).reshape(B, m)
        self.x_k_ = np.minimum(np.maximum(x0, self.x_l_), self.x_u_)  # shift_starting_point, src/SQPTNLP.cpp:140-153
        self.multiplier_cons_ = np.tile(np.asarray(lam_start, dtype=np.float64), (B, 1)).reshape(B, m)
        self.multiplier_vars_ = np.zeros((B, n))
        if hasattr(nlp, "Eval_all"):  # fused device evaluation (nl_reader.DeviceNLP): one launch for everything
            self.obj_value_, self.c_k_, self.grad_f_, self.jac_val_, self.hess_val_ = nlp.Eval_all(self.x_k_, -self.multiplier_cons_)
        else:
            self.obj_value_ = nlp.Eval_f(self.x_k_)
            self.grad_f_ = nlp.Eval_gradient(self.x_k_)
            self.c_k_ = nlp.Eval_constraints(self.x_k_)
            self.hess_val_ = nlp.Eval_Hessian(self.x_k_, -self.multiplier_cons_)
            self.jac_val_ = nlp.Eval_Jacobian(self.x_k_)
        self.cons_type_ = classify_single_constraint(self.c_l_, self.c_u_)           # :467
        self.bound_cons_type_ = classify_single_constraint(self.x_l_, self.x_u_)
        self.infea_measure_ = self.cal_infea(self.c_k_)                              # :472
        self.exitflag_ = np.full(B, int(Exitflag.UNKNOWN), dtype=np.int32)
        self.lp_failed_ = np.zeros(B, dtype=bool)
        self.iter_ = np.zeros(B, dtype=np.int64)
        self.qp_iter_ = np.zeros(B, dtype=np.int64)
        self.penalty_change_trial_ = np.zeros(B, dtype=np.int64)
        f = lambda: np.zeros(B, dtype=bool)
        self.Update_A, self.Update_H, self.Update_bounds, self.Update_delta = f(), f(), f(), f()
        self.Update_penalty, self.Update_g = f(), f()
        self.p_k_ = np.zeros((B, n))
        self.KKT_error_ = np.full(B, np.inf)
        self.first_ = True
        self.lp_set_ = False
        self.actual_reduction_ = np.zeros(B)
        self.pred_reduction_ = np.zeros(B)
        self.x_trial_ = self.x_k_.copy()
        self.c_trial_ = self.c_k_.copy()
        self.obj_value_trial_ = self.obj_value_.copy()
        self.infea_measure_trial_ = self.infea_measure_.copy()
        self.infea_measure_model_ = np.zeros(B)
        self.norm_p_k_ = np.zeros(B)

    # ---- triplet wrappers
    def _jac(self):
        return SpTripletMat(self.nlp_.J_row1, self.nlp_.J_col1, self.jac_val_, self.nCon_, self.nVar_, False)

    def _hess(self):
        return SpTripletMat(self.nlp_.H_row1, self.nlp_.H_col1, self.hess_val_, self.nVar_, self.nVar_, True)

    # ---- src/Algorithm.cpp:577-602 (c part; the x part is only used under NEW_FORMULATION)
    def cal_infea(self, c):
        below = np.where(c < self.c_l_, self.c_l_ - c, 0.0)
        above = np.where((c >= self.c_l_) & (c > self.c_u_), c - self.c_u_, 0.0)
        s = np.zeros(c.shape[0])
        for i in range(c.shape[1]):  # index order of the reference loop
            s = s + below[:, i] + above[:, i]
        return s

    # ---- src/Algorithm.cpp:645-697
    def setupQP(self, active):
        qp = self.myQP_
        if self.first_:
            qp.set_A(self._jac())
            qp.set_H(self._hess())
            qp.set_bounds(self.delta_, self.x_l_, self.x_u_, self.x_k_, self.c_l_, self.c_u_, self.c_k_)
            qp.set_g(self.grad_f_, self.rho_)
            self.first_ = False
            return
        a = active
        # no Update_* flag raised since the instance's last solve: the reference throws QP_UNCHANGED (src/Algorithm.cpp:651-670),
        # which nothing catches; here the instance ends with its own exit flag instead of re-solving the same QP until iter_max
        unchanged = a & ~(self.Update_A | self.Update_H | self.Update_bounds | self.Update_delta | self.Update_penalty | self.Update_g)
        if unchanged.any():
            self.exitflag_[unchanged] = int(Exitflag.QP_UNCHANGED)
            a = a & ~unchanged
        if (self.Update_A & a).any():
            qp.update_A(self._jac())
        if (self.Update_H & a).any():
            qp.update_H(self._hess())
        if (self.Update_bounds & a).any():
            qp.update_bounds(self.delta_, self.x_l_, self.x_u_, self.x_k_, self.c_l_, self.c_u_, self.c_k_)
        elif (self.Update_delta & a).any():
            qp.update_delta(self.delta_, self.x_l_, self.x_u_, self.x_k_)
        if (self.Update_penalty & a).any():
            qp.update_penalty(self.rho_)
        if (self.Update_g & a).any():
            qp.update_grad(self.grad_f_)
        for fl in (self.Update_A, self.Update_H, self.Update_bounds, self.Update_delta, self.Update_penalty, self.Update_g):
            fl[a] = False

    def _solveQP(self, mask):
        """myQP_->solveQP with QP_NOT_OPTIMAL handling (:64-72): failed instances leave the loop with the QP status."""
        st = Stats()
        try:
            ok = self.myQP_.solveQP(st, self.options_, active_mask=mask.astype(np.uint8))
        except QP_NOT_OPTIMAL:  # batch == 1 keeps the reference's exception behaviour
            ok = np.zeros(self.batch, dtype=bool)
        if st.qp_iter.shape == (self.batch,):
            self.qp_iter_[mask] += st.qp_iter[mask]
        status = self.myQP_.get_status()
        failed = mask & ~(np.asarray(ok, dtype=bool) & (status == int(Exitflag.QP_OPTIMAL)))
        self.exitflag_[failed] = np.where(status[failed] == int(Exitflag.QP_OPTIMAL),
                                          int(Exitflag.QPERROR_INTERNAL_ERROR), status[failed])
        if self.dump_dir_ and failed.any() and self.dump_left_ > 0:
            self._dump_failed(np.where(failed)[0])
        return mask & ~failed

    def _dump_failed(self, idx):
        """src/Algorithm.cpp:69: dump the QP that could not be solved (QORE layout, readable by qp_dump.read_qore_log and by
        the reference's replay driver test/QPsolvers_testers.cpp)."""
        import os
        from . import qp_dump
        si = self.myQP_.solverInterface_
        if not hasattr(si, "getA"):
            return
        os.makedirs(self.dump_dir_, exist_ok=True)
        name = getattr(self.nlp_, "name", type(self.nlp_).__name__)
        if hasattr(si, "get_primal_stacked"):  # QORE-layout backend: its own WriteQPDataToFile writes exactly this file
            for b in idx[: self.dump_left_]:
                si.WriteQPDataToFile(os.path.join(self.dump_dir_, "QORE_%s_inst%dqpdata.log" % (name, int(b))), int(b))
                self.dump_left_ -= 1
            return
        A, Hm = si.getA(), si.getH()
        lb, ub, lbA, ubA, g = si.getLb(), si.getUb(), si.getLbA(), si.getUbA(), si.getG()
        for b in idx[: self.dump_left_]:
            q = dict(nV=si.nV_, nC=si.nC_, lb=lb[b], ub=ub[b], lbA=lbA[b], ubA=ubA[b], g=g[b],
                     A_colptr=A["ColIndex"], A_rowidx=A["RowIndex"], A_val=A["MatVal"][b],
                     H_colptr=Hm["ColIndex"], H_rowidx=Hm["RowIndex"], H_val=Hm["MatVal"][b])
            qp_dump.write_qore_log(os.path.join(self.dump_dir_, "QORE_%s_inst%dqpdata.log" % (name, int(b))), q)
            self.dump_left_ -= 1

    # ---- src/Algorithm.cpp:414-429
    def get_trial_point_info(self, mask):
        self.x_trial_[mask] = (self.x_k_ + self.p_k_)[mask]
        if hasattr(self.nlp_, "Eval_f_c"):
            f_t, c_t = self.nlp_.Eval_f_c(self.x_trial_)
        else:
            f_t, c_t = self.nlp_.Eval_f(self.x_trial_), self.nlp_.Eval_constraints(self.x_trial_)
        self.obj_value_trial_[mask] = f_t[mask]
        self.c_trial_[mask] = c_t[mask]
        self.infea_measure_trial_[mask] = self.cal_infea(self.c_trial_)[mask]

    def get_multipliers(self, mask):  # :618-630 (qpOASES-order backend)
        self.multiplier_cons_[mask] = self.myQP_.get_multipliers_constr()[mask]
        self.multiplier_vars_[mask] = self.myQP_.get_multipliers_bounds()[:, :self.nVar_][mask]

    # ---- src/Algorithm.cpp:886-1028
    def update_penalty_parameter(self, active):
        o = self.options_
        if not o.penalty_update:
            return
        n = self.nVar_
        self.infea_measure_model_[active] = self.myQP_.get_infea_measure_model()[active]
        need = active & (self.infea_measure_model_ > o.penalty_update_tol)
        if not need.any():
            return
        infea_model_tmp = self.infea_measure_model_.copy()
        rho_trial = self.rho_.copy()
        # setupLP (:700-704) + solveLP
        lp = self.myLP_
        lp.set_bounds(self.delta_, self.x_l_, self.x_u_, self.x_k_, self.c_l_, self.c_u_, self.c_k_)
        lp.set_g(None, self.rho_)
        lp.set_A(self._jac())
        st = Stats()
        try:
            lp.solveLP(st, active_mask=need.astype(np.uint8))
        except LP_NOT_OPTIMAL:
            pass
        if st.qp_iter.shape == (self.batch,):
            self.qp_iter_[need] += st.qp_iter[need]
        lp_status = lp.get_status()
        lp_fail = need & (lp_status != int(Exitflag.QP_OPTIMAL))
        self.exitflag_[lp_fail] = lp_status[lp_fail]
        self.lp_failed_ |= lp_fail  # LP_NOT_OPTIMAL leaves Optimize (:900-906)
        need = need & ~lp_fail
        infea_infty = lp.get_infea_measure_model()
        feasible_lp = infea_infty <= o.penalty_update_tol
        # the two while loops of :914-939 and :941-972, executed in lock step over the instances that still iterate
        while True:
            cont_a = need & feasible_lp & (self.infea_measure_model_ > o.penalty_update_tol) & (rho_trial < o.rho_max)
            cont_b = need & ~feasible_lp & ((self.infea_measure_ - self.infea_measure_model_) <
                                            self.eps1_ * (self.infea_measure_ - infea_infty)) & \
                (self.penalty_change_trial_ < o.penalty_iter_max) & (rho_trial < o.rho_max)
            go = (cont_a | cont_b) & (self.exitflag_ == int(Exitflag.UNKNOWN))
            if not go.any():
                break
            rho_trial[go] = np.minimum(o.rho_max, rho_trial[go] * o.increase_parm)
            self.penalty_change_trial_[go] += 1
            self.myQP_.update_penalty(rho_trial)  # rho_trial == rho_ for every instance that never entered the loop
            ok = self._solveQP(go)
            self.infea_measure_model_[ok] = self.myQP_.get_infea_measure_model()[ok]
            # QP_NOT_OPTIMAL inside the loop only leaves the loop (:932-935, :958-961): the acceptance test below then sees the
            # objective of an unsolved QP (INFTY) and takes its failure branch, and Optimize runs the rest of the iteration
        changed = need & (rho_trial > self.rho_)
        if changed.any():
            qp_obj = self.myQP_.get_objective()
            succ = changed & (rho_trial * self.infea_measure_ - qp_obj >=
                              o.eps2 * rho_trial * (self.infea_measure_ - self.infea_measure_model_))
            fail = changed & ~succ
            if succ.any():
                self.eps1_[succ] += (1 - self.eps1_[succ]) * o.eps1_change_parm
                self.p_k_[succ] = self.myQP_.get_optimal_solution()[:, :n][succ]
                self.rho_[succ] = rho_trial[succ]
                self.get_trial_point_info(succ)
                P1_x = self.obj_value_ + self.rho_ * self.infea_measure_
                P1_t = self.obj_value_trial_ + self.rho_ * self.infea_measure_trial_
                self.actual_reduction_[succ] = (P1_x - P1_t)[succ]
                self.pred_reduction_[succ] = (self.rho_ * self.infea_measure_ - qp_obj)[succ]
            if fail.any():
                self.infea_measure_model_[fail] = infea_model_tmp[fail]
                self.Update_penalty[fail] = True
                # the backend still holds rho_trial in g: the flag makes setupQP restore rho_ next iteration (:1003-1006)

    # ---- src/Algorithm.cpp:722-801
    def ratio_test(self, active, qp_obj=None):
        o = self.options_
        if qp_obj is None:
            qp_obj = self.myQP_.get_objective()
        P1_x = self.obj_value_ + self.rho_ * self.infea_measure_
        P1_t = self.obj_value_trial_ + self.rho_ * self.infea_measure_trial_
        self.actual_reduction_[active] = (P1_x - P1_t)[active]
        self.pred_reduction_[active] = (self.rho_ * self.infea_measure_ - qp_obj)[active]
        acc = active & (self.actual_reduction_ >= o.eta_s * self.pred_reduction_) & (self.actual_reduction_ >= -o.tol)
        if acc.any():
            self.infea_measure_[acc] = self.infea_measure_trial_[acc]
            self.obj_value_[acc] = self.obj_value_trial_[acc]
            self.x_k_[acc] = self.x_trial_[acc]
            self.c_k_[acc] = self.c_trial_[acc]
            self.get_multipliers(acc)
            if hasattr(self.nlp_, "Eval_all"):
                _, _, g_n, j_n, h_n = self.nlp_.Eval_all(self.x_k_, -self.multiplier_cons_)
            else:
                g_n, j_n = self.nlp_.Eval_gradient(self.x_k_), self.nlp_.Eval_Jacobian(self.x_k_)
                h_n = self.nlp_.Eval_Hessian(self.x_k_, -self.multiplier_cons_)
            self.grad_f_[acc] = g_n[acc]
            self.jac_val_[acc] = j_n[acc]
            self.hess_val_[acc] = h_n[acc]
            self.Update_A[acc] = self.Update_H[acc] = self.Update_bounds[acc] = self.Update_g[acc] = True
        return acc

    # ---- src/Algorithm.cpp:1140-1211 (off by default, src/Options.cpp:26; the reference marks it "FIXME: check correctness")
    def second_order_correction(self, rej):
        """For the instances whose step was rejected: solve the QP again around the trial point with the gradient H_k p_k + g_k,
        add its solution s_k to p_k and repeat the ratio test; restore p_k and the QP data if the corrected step is rejected too.
        Instances outside `rej` keep their QP data (the mixed arrays hold their current values)."""
        rej = rej & (self.exitflag_ == int(Exitflag.UNKNOWN))
        if not rej.any():
            return
        n, B = self.nVar_, self.batch
        p_tmp = self.p_k_.copy()
        qp_obj_tmp = self.myQP_.get_objective()
        # Hp = H_k p_k + grad_f: symmetric-half triplet product in storage order (src/SpTripletMat.cpp:237-258)
        Hp = np.zeros((B, n))
        for k in range(len(self.nlp_.H_row1)):
            i, j = self.nlp_.H_row1[k] - 1, self.nlp_.H_col1[k] - 1
            Hp[:, i] += self.hess_val_[:, k] * self.p_k_[:, j]
            if i != j:
                Hp[:, j] += self.hess_val_[:, k] * self.p_k_[:, i]
        Hp = Hp + self.grad_f_
        m_ = rej[:, None]
        self.myQP_.update_grad(np.where(m_, Hp, self.grad_f_))
        self.myQP_.update_bounds(self.delta_, self.x_l_, self.x_u_, np.where(m_, self.x_trial_, self.x_k_), self.c_l_, self.c_u_,
                                 np.where(m_, self.c_trial_, self.c_k_))
        ok = self._solveQP(rej)
        s_k = self.myQP_.get_optimal_solution()[:, :n]
        qp_obj_soc = self.myQP_.get_objective() + (qp_obj_tmp - self.rho_ * self.infea_measure_model_)
        self.p_k_[ok] = (self.p_k_ + s_k)[ok]
        self.get_trial_point_info(ok)
        # ratio_test takes pred_reduction_ from get_obj_QP(), the raw objective of the SOC QP just solved (src/Algorithm.cpp:728);
        # qp_obj_ (= qp_obj_soc here) is bookkeeping the reference never reads in a test
        del qp_obj_soc
        acc2 = self.ratio_test(ok, qp_obj=self.myQP_.get_objective())
        took = ok & acc2  # update_radius reads p_k_->getInfNorm() (:822): after an accepted correction the norm of p_k + s_k
        if took.any() and self.nVar_:
            self.norm_p_k_[took] = np.abs(self.p_k_[took]).max(axis=1)
        still = rej & ~acc2
        self.p_k_[still] = p_tmp[still]
        self.myQP_.update_grad(self.grad_f_)
        self.myQP_.update_bounds(self.delta_, self.x_l_, self.x_u_, self.x_k_, self.c_l_, self.c_u_, self.c_k_)

    # ---- src/Algorithm.cpp:170-411
    def check_optimality(self, active):
        o, n, m = self.options_, self.nVar_, self.nCon_
        self.get_multipliers(active)
        mv, mc = self.multiplier_vars_, self.multiplier_cons_
        B = self.batch
        primal = self.infea_measure_.copy()
        dual, compl = np.zeros(B), np.zeros(B)
        bt, ct = self.bound_cons_type_, self.cons_type_
        for i in range(n):
            dual = dual + np.where(bt[:, i] == BOUNDED_ABOVE, np.maximum(mv[:, i], 0.0), 0.0) \
                + np.where(bt[:, i] == BOUNDED_BELOW, -np.minimum(mv[:, i], 0.0), 0.0)
        for i in range(m):
            dual = dual + np.where(ct[:, i] == BOUNDED_ABOVE, np.maximum(mc[:, i], 0.0), 0.0) \
                + np.where(ct[:, i] == BOUNDED_BELOW, -np.minimum(mc[:, i], 0.0), 0.0)
        for i in range(m):
            compl = compl + np.where(ct[:, i] == BOUNDED_ABOVE, np.abs(mc[:, i] * (self.c_u_[:, i] - self.c_k_[:, i])), 0.0) \
                + np.where(ct[:, i] == BOUNDED_BELOW, np.abs(mc[:, i] * (self.c_k_[:, i] - self.c_l_[:, i])), 0.0) \
                + np.where(ct[:, i] == UNBOUNDED, np.abs(mc[:, i]), 0.0)
        for i in range(n):
            compl = compl + np.where(bt[:, i] == BOUNDED_ABOVE, np.abs(mv[:, i] * (self.x_u_[:, i] - self.x_k_[:, i])), 0.0) \
                + np.where(bt[:, i] == BOUNDED_BELOW, np.abs(mv[:, i] * (self.x_k_[:, i] - self.x_l_[:, i])), 0.0) \
                + np.where(bt[:, i] == UNBOUNDED, np.abs(mv[:, i]), 0.0)
        # stationarity: || J' y_c + y_b - grad_f ||_1 with the triplet SpMTV of src/SpTripletMat.cpp:311-323
        diff = np.zeros((B, n))
        for k in range(len(self.nlp_.J_row1)):
            diff[:, self.nlp_.J_col1[k] - 1] += self.jac_val_[:, k] * mc[:, self.nlp_.J_row1[k] - 1]
        diff = diff + mv - self.grad_f_
        stat = np.zeros(B)
        for i in range(n):
            stat = stat + np.abs(diff[:, i])
        kkt = dual + primal + compl + stat
        self.KKT_error_[active] = kkt[active]
        opt = active & (primal < o.opt_prim_fea_tol) & (dual < o.opt_dual_fea_tol) & (compl < o.opt_compl_tol) & (stat < o.opt_stat_tol)
        self.exitflag_[opt] = int(Exitflag.OPTIMAL)

    # ---- src/Algorithm.cpp:820-849
    def update_radius(self, active):
        o = self.options_
        shrink = active & (self.actual_reduction_ < o.eta_c * self.pred_reduction_)
        norm_p = self.norm_p_k_  # ||p_k||_inf of the step taken (:822): the corrected step after an accepted second-order correction
        grow = active & ~shrink & (self.actual_reduction_ > o.eta_e * self.pred_reduction_) & (o.tol > np.abs(self.delta_ - norm_p))
        self.delta_[shrink] = o.gamma_c * self.delta_[shrink]
        self.delta_[grow] = np.minimum(o.gamma_e * self.delta_[grow], o.delta_max)
        self.Update_delta[shrink | grow] = True
        small = active & (self.delta_ < o.delta_min)
        if small.any():
            self.exitflag_[small] = int(Exitflag.TRUST_REGION_TOO_SMALL)
            self.check_optimality(small)  # :149-152

    # ---- src/Algorithm.cpp:55-168
    def Optimize(self):
        o, n = self.options_, self.nVar_
        UNK = int(Exitflag.UNKNOWN)
        while True:
            active = (self.iter_ < o.iter_max) & (self.exitflag_ == UNK)
            if not active.any():
                break
            self.setupQP(active)
            active = active & (self.exitflag_ == UNK)
            if not active.any():
                break
            active = self._solveQP(active)
            self.p_k_[active] = self.myQP_.get_optimal_solution()[:, :n][active]   # get_search_direction :609
            self.update_penalty_parameter(active)
            active = active & ~self.lp_failed_
            self.norm_p_k_ = np.abs(self.p_k_).max(axis=1) if self.nVar_ else np.zeros(self.batch)
            self.get_trial_point_info(active)
            acc = self.ratio_test(active)
            if o.second_order_correction:
                self.second_order_correction(active & ~acc)  # (instances whose exit flag is already set are skipped inside)
            self.iter_[active] += 1
            self.check_optimality(active)
            still = active & (self.exitflag_ == UNK)
            self.update_radius(still)
        self.exitflag_[(self.iter_ == o.iter_max) & (self.exitflag_ == UNK)] = int(Exitflag.EXCEED_MAX_ITER)
        return SQPResult(x=self.x_k_.copy(), obj=self.obj_value_.copy(), exitflag=self.exitflag_.copy(), iters=self.iter_.copy(),
                         qp_iter=self.qp_iter_.copy(), KKT_error=self.KKT_error_.copy(), rho=self.rho_.copy(), delta=self.delta_.copy())
