"""Device-resident batched Sl1QP driver.

Same algorithm and same arithmetic as `sqp_driver.BatchedSQP` (the host mirror of src/Algorithm.cpp), but the iterates, the
trust-region radius, the penalty parameter, the flags and every intermediate stay on the GPU: the per-instance steps are phases
of one CUDA kernel (csrc/sqp_outer.cu, one thread per instance), the NLP is evaluated by the NVRTC-compiled kernels of
`DeviceNLP`, QP data is built by the qphandler kernels and the QP/LP solves take their instance mask from device memory.  The
host sequences launches and reads eight counters per outer iteration.  torch is used for device memory only.

By default the backend's init / hotstart decision (src/qpOASESInterface.cpp:141-211, 817-833) is made per instance inside the solve
kernel, so every instance of the batch follows the reference's single-instance semantics: tests/test_gpu_sqp.py compares it bitwise
with the C oracle of the loop (oracle/oracle_sqp.c), and, with per_instance_modes=False, with BatchedSQP.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from .qp_handler import QPhandler
from .sqp_driver import SQPResult, classify_single_constraint
from .sqp_types import Exitflag, Options, QPType, SpTripletMat

PH_FLAGS, PH_AFTER_QP, PH_LP_AFTER, PH_PEN_CHECK, PH_PEN_AFTER, PH_PEN_FINAL, PH_TRIAL, PH_RATIO, PH_FINISH, PH_FINAL = range(10)
PH_SOC_PREP, PH_SOC_AFTER, PH_SOC_RATIO, PH_INIT = 10, 11, 12, 13
UP_A, UP_H, UP_BOUNDS, UP_DELTA, UP_PENALTY, UP_G = 1, 2, 4, 8, 16, 32

_P, _D, _I = C.c_void_p, C.c_double, C.c_int


class SqpState(C.Structure):
    """sqpb200_sqp_state of include/sqpb200.h"""
    _fields_ = ([(k, _I) for k in ("B", "n", "m", "zJ", "zH", "iter_max", "penalty_update", "penalty_iter_max", "clear_flags")] +
                [(k, _D) for k in ("eta_c", "eta_s", "eta_e", "gamma_c", "gamma_e", "delta_min", "delta_max", "tol", "penalty_update_tol",
                                   "rho_max", "increase_parm", "eps1_change_parm", "eps2", "opt_prim_fea_tol", "opt_dual_fea_tol",
                                   "opt_compl_tol", "opt_stat_tol")] +
                [(k, _P) for k in ("J_row1", "J_col1", "x_l", "x_u", "c_l", "c_u", "bound_type", "cons_type",
                                   "x_k", "c_k", "f_k", "grad", "jac", "hess", "lam_c", "lam_x", "neg_lam",
                                   "delta", "rho", "eps1", "infea", "p_k", "x_trial", "c_trial", "f_trial", "infea_trial", "infea_model",
                                   "infea_model_tmp", "rho_trial", "infea_infty", "actual_red", "pred_red", "kkt_err",
                                   "g_new", "j_new", "h_new", "scratch", "exitflag", "iter", "pen_trial", "qp_iter",
                                   "active", "need", "go", "acc", "upd", "feasible_lp",
                                   "qp_x", "qp_y", "qp_obj", "qp_kkt", "lp_x", "qp_status", "qp_iters", "lp_status", "lp_iters", "counters",
                                   "H_row1", "H_col1", "soc_g", "soc_x", "soc_c", "p_tmp", "qp_obj_tmp", "qp_obj_soc", "norm_p", "rej",
                                   "qp_inst", "lp_inst")] +
                [(k, _D) for k in ("delta0", "rho0", "eps10")])


class DeviceBatchedSQP:
    def __init__(self, nlp, x0=None, options: Options = None, device=0, per_instance_modes=True):
        """per_instance_modes: the backend's init/hotstart decision is made per instance, as the reference does for its single
        instance (and as oracle/oracle_sqp.c does); False: per handle, which is what the numpy mirror BatchedSQP does."""
        import torch
        if not hasattr(nlp, "eval_device"):
            raise TypeError("DeviceBatchedSQP needs a DeviceNLP (device-side evaluation)")
        self.torch, self.nlp_, self.L = torch, nlp, capi.lib()
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        o = self.options_ = options if options is not None else Options()
        info = self.info = nlp.Get_nlp_info()
        n, m = self.nVar_, self.nCon_ = info.nVar, info.nCon
        x_start, lam_start = nlp.Get_starting_point()
        x0 = np.atleast_2d(np.asarray(x_start if x0 is None else x0, dtype=np.float64))
        B = self.batch = x0.shape[0]
        zJ, zH = len(nlp.J_row1), len(nlp.H_row1)
        self.myQP_ = QPhandler(info, QPType.QP, o, batch=B, device=device, refresh_ubA=True)  # src/Algorithm.cpp:561-562
        self.myLP_ = QPhandler(info, QPType.LP, o, batch=B, device=device, refresh_ubA=True)
        f64 = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=self.dev)
        u8 = lambda: torch.zeros(B, dtype=torch.uint8, device=self.dev)
        up = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(self.dev)
        xl, xu, cl, cu = nlp.Get_bounds_info()
        # bounds, constraint classes and multipliers are the same for every instance: upload one row, replicate on the device
        rep = lambda a, dt=torch.float64: up(np.asarray(a).reshape(1, -1), dt).repeat(B, 1).contiguous()
        T = self.T = {}
        T["x_l"], T["x_u"], T["c_l"], T["c_u"] = rep(xl), rep(xu), rep(cl), rep(cu)
        T["bound_type"] = rep(classify_single_constraint(np.atleast_2d(xl), np.atleast_2d(xu)), torch.int32)
        T["cons_type"] = rep(classify_single_constraint(np.atleast_2d(cl).reshape(1, m), np.atleast_2d(cu).reshape(1, m)), torch.int32)
        T["J_row1"], T["J_col1"] = up(nlp.J_row1, torch.int32), up(nlp.J_col1, torch.int32)
        T["H_row1"], T["H_col1"] = up(nlp.H_row1, torch.int32), up(nlp.H_col1, torch.int32)
        T["x_k"] = torch.minimum(torch.maximum(up(x0), T["x_l"]), T["x_u"])  # shift_starting_point, src/SQPTNLP.cpp:140-153
        T["lam_c"] = rep(np.asarray(lam_start, dtype=np.float64).reshape(1, m))
        T["neg_lam"] = -T["lam_c"]
        for k, shp in (("c_k", (B, m)), ("f_k", (B,)), ("grad", (B, n)), ("jac", (B, zJ)), ("hess", (B, zH)), ("lam_x", (B, n)),
                       ("infea", (B,)), ("p_k", (B, n)), ("x_trial", (B, n)), ("c_trial", (B, m)), ("f_trial", (B,)), ("infea_trial", (B,)),
                       ("infea_model", (B,)), ("infea_model_tmp", (B,)), ("rho_trial", (B,)), ("infea_infty", (B,)), ("actual_red", (B,)),
                       ("pred_red", (B,)), ("g_new", (B, n)), ("j_new", (B, zJ)), ("h_new", (B, zH)), ("scratch", (B, max(n, 1))),
                       ("f_tmp", (B,)), ("c_tmp", (B, m)), ("soc_g", (B, n)), ("soc_x", (B, n)), ("soc_c", (B, m)), ("p_tmp", (B, n)),
                       ("qp_obj_tmp", (B,)), ("qp_obj_soc", (B,)), ("norm_p", (B,))):
            T[k] = f64(*shp)
        T["delta"] = torch.full((B,), float(o.delta), dtype=torch.float64, device=self.dev)
        T["rho"] = torch.full((B,), float(o.rho), dtype=torch.float64, device=self.dev)
        T["eps1"] = torch.full((B,), float(o.eps1), dtype=torch.float64, device=self.dev)
        T["kkt_err"] = torch.full((B,), float("inf"), dtype=torch.float64, device=self.dev)
        T["exitflag"] = torch.full((B,), int(Exitflag.UNKNOWN), dtype=torch.int32, device=self.dev)
        T["iter"] = torch.zeros(B, dtype=torch.int32, device=self.dev)
        T["pen_trial"] = torch.zeros(B, dtype=torch.int32, device=self.dev)
        T["qp_iter"] = torch.zeros(B, dtype=torch.int64, device=self.dev)
        for k in ("active", "need", "go", "acc", "upd", "feasible_lp", "rej"):
            T[k] = u8()
        T["counters"] = torch.zeros(8, dtype=torch.int32, device=self.dev)
        self.per_instance_modes = bool(per_instance_modes)
        if self.per_instance_modes:
            for k in ("qp_inst", "lp_inst"):
                T[k] = torch.zeros((B, 8), dtype=torch.int8, device=self.dev)
                T[k][:, 2:4] = -1  # matrix status undefined (get_Matrix_change_status, src/qpOASESInterface.cpp:817-833)
        # initialization(), src/Algorithm.cpp:438-472: f, c, grad, Jacobian, Hessian at the (shifted) start in one launch
        nlp.eval_device(1, B, T["x_k"], T["neg_lam"], T["f_k"], T["c_k"], T["grad"], T["jac"], T["hess"])
        S = self.S = SqpState()
        S.B, S.n, S.m, S.zJ, S.zH = B, n, m, zJ, zH
        S.iter_max, S.penalty_update, S.penalty_iter_max, S.clear_flags = o.iter_max, int(o.penalty_update), o.penalty_iter_max, 0
        for k in ("eta_c", "eta_s", "eta_e", "gamma_c", "gamma_e", "delta_min", "delta_max", "tol", "penalty_update_tol", "rho_max",
                  "increase_parm", "eps1_change_parm", "eps2", "opt_prim_fea_tol", "opt_dual_fea_tol", "opt_compl_tol", "opt_stat_tol"):
            setattr(S, k, float(getattr(o, k)))
        for k, _ in SqpState._fields_:
            if k in T:
                setattr(S, k, T[k].data_ptr())
        bq, bl = (C.c_void_p * 6)(), (C.c_void_p * 6)()
        self.L.sqpb200_device_buffers(self.myQP_.solverInterface_.h, bq)
        self.L.sqpb200_device_buffers(self.myLP_.solverInterface_.h, bl)
        S.qp_x, S.qp_y, S.qp_obj, S.qp_status, S.qp_iters, S.qp_kkt = bq[0], bq[1], bq[2], bq[3], bq[4], bq[5]
        S.lp_x, S.lp_status, S.lp_iters = bl[0], bl[3], bl[4]
        self._counters = (C.c_int * 8)()
        self.first_ = True
        self.launches = 0
        S.delta0, S.rho0, S.eps10 = float(o.delta), float(o.rho), float(o.eps1)
        self._lam_start = np.asarray(lam_start, dtype=np.float64).reshape(1, m)
        # infea_measure_ of the starting point (:472) and the per-instance algorithm state
        self._phase(PH_INIT)

    def reset(self, x0):
        """Algorithm::initialization (src/Algorithm.cpp:438-472) for a new batch of `batch` starting points on the same object:
        handles, device buffers and the compiled NLP are kept, so the per-batch cost is one upload, one evaluation and one
        state-reset launch."""
        torch, T, B = self.torch, self.T, self.batch
        x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)))
        if x0.shape != (B, self.nVar_):
            raise ValueError("reset() needs %d starting points of dimension %d" % (B, self.nVar_))
        T["x_k"].copy_(torch.from_numpy(x0), non_blocking=False)
        torch.minimum(torch.maximum(T["x_k"], T["x_l"]), T["x_u"], out=T["x_k"])  # shift_starting_point
        T["lam_c"].copy_(torch.from_numpy(self._lam_start).to(self.dev).expand(B, -1))
        torch.neg(T["lam_c"], out=T["neg_lam"])
        self.nlp_.eval_device(1, B, T["x_k"], T["neg_lam"], T["f_k"], T["c_k"], T["grad"], T["jac"], T["hess"])
        self.S.clear_flags = 0
        self._phase(PH_INIT)
        for hd in (self.myQP_, self.myLP_):
            self.L.sqpb200_reset(hd.solverInterface_.h)
        self.first_ = True

    # ---- helpers
    def _phase(self, phase, read=False):
        rc = self.L.sqpb200_sqp_phase(C.byref(self.S), phase, self._counters if read else None, None)
        if rc != 0:
            raise capi.SqpB200Error("sqpb200_sqp_phase(%d) failed: %d" % (phase, rc))
        self.launches += 1
        return list(self._counters) if read else None

    def _jac(self):
        return SpTripletMat(self.nlp_.J_row1, self.nlp_.J_col1, self.T["jac"], self.nCon_, self.nVar_, False)

    def _hess(self):
        return SpTripletMat(self.nlp_.H_row1, self.nlp_.H_col1, self.T["hess"], self.nVar_, self.nVar_, True)

    def _solve(self, handler, qptype, mask):
        si = handler.solverInterface_
        if self.per_instance_modes:
            inst = self.T["qp_inst" if handler is self.myQP_ else "lp_inst"]
            rc = self.L.sqpb200_solve_per_instance(si.h, int(qptype), 0, C.c_void_p(mask.data_ptr()), C.c_void_p(inst.data_ptr()))
        else:
            rc = self.L.sqpb200_solve_device_mask(si.h, int(qptype), 0, C.c_void_p(mask.data_ptr()))
        if rc != 0:
            raise capi.SqpB200Error("batched solve failed (%d): %s" % (rc, self.L.sqpb200_last_error(si.h).decode()))
        si._kkt = None

    # ---- src/Algorithm.cpp:645-697
    def setupQP(self, bits):
        T, qp = self.T, self.myQP_
        if self.first_:
            qp.set_A(self._jac())
            qp.set_H(self._hess())
            qp.set_bounds(T["delta"], T["x_l"], T["x_u"], T["x_k"], T["c_l"], T["c_u"], T["c_k"])
            qp.set_g(T["grad"], T["rho"])
            self.first_ = False
            self.S.clear_flags = 1
            return
        if bits & UP_A:
            qp.update_A(self._jac())
        if bits & UP_H:
            qp.update_H(self._hess())
        if bits & UP_BOUNDS:
            qp.update_bounds(T["delta"], T["x_l"], T["x_u"], T["x_k"], T["c_l"], T["c_u"], T["c_k"])
        elif bits & UP_DELTA:
            qp.update_delta(T["delta"], T["x_l"], T["x_u"], T["x_k"])
        if bits & UP_PENALTY:
            qp.update_penalty(T["rho"])
        if bits & UP_G:
            qp.update_grad(T["grad"])

    # ---- src/Algorithm.cpp:886-1028
    def update_penalty_parameter(self):
        T, lp = self.T, self.myLP_
        lp.set_bounds(T["delta"], T["x_l"], T["x_u"], T["x_k"], T["c_l"], T["c_u"], T["c_k"])  # setupLP :700-704
        lp.set_g(None, T["rho"])
        lp.set_A(self._jac())
        self._solve(lp, QPType.LP, T["need"])
        self._phase(PH_LP_AFTER)
        while True:
            go = self._phase(PH_PEN_CHECK, read=True)[3]
            if go == 0:
                break
            self.myQP_.update_penalty(T["rho_trial"])
            self._solve(self.myQP_, QPType.QP, T["go"])
            self._phase(PH_PEN_AFTER)
        self._phase(PH_PEN_FINAL)

    # ---- src/Algorithm.cpp:1140-1211 (opt-in)
    def second_order_correction(self):
        T, qp, B = self.T, self.myQP_, self.batch
        if self._phase(PH_SOC_PREP, read=True)[5] == 0:
            return
        qp.update_grad(T["soc_g"])
        qp.update_bounds(T["delta"], T["x_l"], T["x_u"], T["soc_x"], T["c_l"], T["c_u"], T["soc_c"])
        self._solve(qp, QPType.QP, T["rej"])
        self._phase(PH_SOC_AFTER)
        self.nlp_.eval_device(0, B, T["x_trial"], None, T["f_trial"], T["c_trial"])
        self._phase(PH_SOC_RATIO)
        qp.update_grad(T["grad"])
        qp.update_bounds(T["delta"], T["x_l"], T["x_u"], T["x_k"], T["c_l"], T["c_u"], T["c_k"])

    # ---- src/Algorithm.cpp:55-168
    def Optimize(self, host_sequenced=False):
        """The whole loop runs behind one C call (sqpb200_sqp_optimize: no interpreter between the launches);
        host_sequenced=True sequences the same launches from Python (the two are compared bitwise in tests/test_gpu_sqp.py)."""
        if host_sequenced:
            return self._optimize_py()
        T = self.T
        if self.first_:  # structures of both backends (first set_A / set_H call of setupQP / setupLP, values follow in the loop)
            self.myQP_.solverInterface_.set_A(SpTripletMat(self.nlp_.J_row1, self.nlp_.J_col1, None, self.nCon_, self.nVar_, False), self.myQP_.I_info_A_)
            self.myQP_.solverInterface_.set_H(SpTripletMat(self.nlp_.H_row1, self.nlp_.H_col1, None, self.nVar_, self.nVar_, True))
        if not self.myLP_.solverInterface_._A_set:
            self.myLP_.solverInterface_.set_A(SpTripletMat(self.nlp_.J_row1, self.nlp_.J_col1, None, self.nCon_, self.nVar_, False), self.myLP_.I_info_A_)
        first, nl = C.c_int(1 if self.first_ else 0), C.c_longlong(0)
        rc = self.L.sqpb200_sqp_optimize(C.byref(self.S), self.myQP_.solverInterface_.h, self.myLP_.solverInterface_.h, self.nlp_.h,
                                         int(bool(self.options_.second_order_correction)), 1, C.byref(first),
                                         C.c_void_p(T["f_tmp"].data_ptr()), C.c_void_p(T["c_tmp"].data_ptr()), C.byref(nl), None)
        if rc != 0:
            raise capi.SqpB200Error("sqpb200_sqp_optimize failed (%d): %s / %s" % (
                rc, self.L.sqpb200_last_error(self.myQP_.solverInterface_.h).decode(), self.L.sqpb200_last_error(self.myLP_.solverInterface_.h).decode()))
        self.first_ = bool(first.value)
        self.launches += int(nl.value)
        return self._result()

    def _result(self):
        T = self.T
        self.torch.cuda.synchronize()
        h = lambda k: T[k].cpu().numpy()
        return SQPResult(x=h("x_k"), obj=h("f_k"), exitflag=h("exitflag"), iters=h("iter").astype(np.int64), qp_iter=h("qp_iter"),
                         KKT_error=h("kkt_err"), rho=h("rho"), delta=h("delta"))

    def _optimize_py(self):
        T, nlp, B = self.T, self.nlp_, self.batch
        while True:
            cnt = self._phase(PH_FLAGS, read=True)
            if cnt[0] == 0:
                break
            self.setupQP(cnt[1])
            self._solve(self.myQP_, QPType.QP, T["active"])
            need = self._phase(PH_AFTER_QP, read=True)[2]
            if need > 0:
                self.update_penalty_parameter()
            self._phase(PH_TRIAL)
            nlp.eval_device(0, B, T["x_trial"], None, T["f_trial"], T["c_trial"])  # get_trial_point_info :414-429
            self._phase(PH_RATIO)
            if self.options_.second_order_correction:
                self.second_order_correction()
            nlp.eval_device(1, B, T["x_k"], T["neg_lam"], T["f_tmp"], T["c_tmp"], T["g_new"], T["j_new"], T["h_new"])
            self._phase(PH_FINISH)
        self._phase(PH_FINAL)
        return self._result()

    def close(self):
        self.myQP_.solverInterface_.close()
        self.myLP_.solverInterface_.close()
