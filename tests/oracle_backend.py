"""TEST INFRASTRUCTURE: a CPU twin of restartsqp_b200.CudaQPInterface built on the oracle (oracle/oracle_qp.c).

Same method set, same batch semantics (per-handle Update_* flags and init/hotstart state machine of
src/qpOASESInterface.cpp:137-284, 817-833), one OracleQP object per instance.  Plugged into the product's QPhandler /
BatchedSQP through their `backend` / `make_handler` hooks so that the host logic can be tested on CPU and the GPU path can
be compared iterate by iterate."""
import numpy as np

from oracle import oracle_py as orc

MS_UNDEFINED, MS_FIXED, MS_VARIED = -1, 0, 1


class OracleQPInterface:
    def __init__(self, nlp_info=None, qptype=2, options=None, batch=1, nV=None, nC=None):
        if nlp_info is not None:
            nC, nV = nlp_info.nCon, nlp_info.nVar + 2 * nlp_info.nCon
        self.nV_, self.nC_, self.batch, self.is_lp = nV, nC, batch, int(qptype) == 1
        self.maxiter = (options.lp_maxiter if self.is_lp else options.qp_maxiter) if options is not None else (100 if self.is_lp else 1000)
        B = batch
        self.g, self.lb, self.ub = np.zeros((B, nV)), np.zeros((B, nV)), np.zeros((B, nV))
        self.lbA, self.ubA = np.zeros((B, nC)), np.zeros((B, nC))
        self.x, self.y, self.obj = np.zeros((B, nV)), np.zeros((B, nV + nC)), np.zeros(B)
        self.status = np.full(B, 25, np.int32)
        self.iters = np.zeros(B, np.int32)
        self.kkt = np.zeros((B, 5))
        self.wb, self.wc = np.zeros((B, nV), np.int32), np.zeros((B, nC), np.int32)
        self.A = self.H = None
        self.qA = self.qH = None
        self.solvers = [orc.OracleQP(nV, nC, max_iter=self.maxiter) for _ in range(B)]
        self.inited = np.zeros(B, bool)
        self.first_solved = False
        self.upd_A = self.upd_H = False
        self.old_ms = self.new_ms = MS_UNDEFINED

    def close(self):
        self.solvers = []

    # ---- data
    def _bvals(self, v, z):
        v = np.asarray(v, dtype=np.float64)
        return np.ascontiguousarray(np.broadcast_to(v, (self.batch, z))).copy()

    def set_A(self, rhs, I_info=None):
        if self.A is None:
            Ap, Ai, Av, Ao = orc.assemble_A(self.nC_, self.nV_, rhs.RowIndex, rhs.ColIndex, np.zeros(len(rhs.RowIndex)),
                                            (I_info.irow, I_info.jcol, I_info.size, I_info.value))
            self.A = dict(p=Ap, i=Ai, order=Ao, init=Av.copy(), zJ=len(rhs.RowIndex))
            self.Av = np.tile(Av, (self.batch, 1))
        vals = self._bvals(rhs.MatVal, self.A["zJ"])
        for b in range(self.batch):
            self.Av[b] = orc.setmatval_A(self.A["order"], vals[b], self.Av[b])
        if self.first_solved:
            self.upd_A = True

    def set_H(self, rhs):
        if self.H is None:
            Hp, Hi, Hv, Ho = orc.assemble_H(self.nV_, rhs.RowIndex, rhs.ColIndex, np.zeros(len(rhs.RowIndex)), rhs.isSymmetric)
            self.H = dict(p=Hp, i=Hi, order=Ho, r=rhs.RowIndex, c=rhs.ColIndex, sym=rhs.isSymmetric)
            self.Hv = np.zeros((self.batch, len(Hi)))
        vals = self._bvals(rhs.MatVal, len(rhs.RowIndex))
        for b in range(self.batch):
            self.Hv[b] = orc.setmatval_H(self.H["r"], self.H["c"], self.H["order"], vals[b], self.Hv[b], self.H["sym"])
        if self.first_solved:
            self.upd_H = True

    def qphandler_bounds(self, mode, n, m, delta, x_l, x_u, x_k, c_l=None, c_u=None, c_k=None):
        z = np.zeros(m)
        for b in range(self.batch):
            orc.qp_bounds(mode, n, m, float(delta[b]), x_l[b], x_u[b], x_k[b], z if c_l is None else c_l[b],
                          z if c_u is None else c_u[b], z if c_k is None else c_k[b], self.lb[b], self.ub[b], self.lbA[b], self.ubA[b])

    def qphandler_g(self, n, m, grad, rho):
        if grad is not None:
            self.g[:, :n] = grad
        if rho is not None:
            self.g[:, n:] = np.asarray(rho)[:, None]

    @staticmethod
    def _put(arr, a0, a1):
        if a1 is None:
            arr[:] = a0        # vector form
        else:
            arr[:, int(a0)] = a1  # (location, value) form

    def set_g(self, a0, a1=None): self._put(self.g, a0, a1)
    def set_lb(self, a0, a1=None): self._put(self.lb, a0, a1)
    def set_ub(self, a0, a1=None): self._put(self.ub, a0, a1)
    def set_lbA(self, a0, a1=None): self._put(self.lbA, a0, a1)
    def set_ubA(self, a0, a1=None): self._put(self.ubA, a0, a1)

    # ---- QORE layout (the twin of the sqpb200_*_csr / *_stacked entry points): row-compressed views, stacked vectors
    def _csr_view(self, rp, ci, ncol):
        """Position map between the row- and the column-compressed storage of one pattern (stable for equal keys)."""
        rp, ci = np.asarray(rp, np.int64), np.asarray(ci, np.int64)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        q2c = np.empty(len(ci), np.int64)
        q2c[np.lexsort((np.arange(len(ci)), rows, ci))] = np.arange(len(ci))
        cp = np.zeros(ncol + 1, np.int32)
        np.add.at(cp, ci + 1, 1)
        cp = np.cumsum(cp).astype(np.int32)
        ri = np.empty(len(ci), np.int32)
        ri[q2c] = rows
        return cp, ri, q2c

    def set_A_csr(self, rhs, I_info=None):
        if self.A is None:
            self.qA = dict(zip(("rp", "ci", "val", "order"), orc.assemble_A_csr(
                self.nC_, self.nV_, rhs.RowIndex, rhs.ColIndex, np.zeros(len(rhs.RowIndex)),
                (I_info.irow, I_info.jcol, I_info.size, I_info.value))))
            self.qA["q2c"] = self._csr_view(self.qA["rp"], self.qA["ci"], self.nV_)[2]
        self.set_A(rhs, I_info)

    def set_H_csr(self, rhs):
        if self.H is None:
            self.qH = dict(zip(("rp", "ci", "val", "order"), orc.assemble_H_csr(
                self.nV_, rhs.RowIndex, rhs.ColIndex, np.zeros(len(rhs.RowIndex)), rhs.isSymmetric)))
            self.qH["q2c"] = self._csr_view(self.qH["rp"], self.qH["ci"], self.nV_)[2]
        self.set_H(rhs)

    def set_csr(self, which, rowptr, colidx, vals):
        cp, ri, q2c = self._csr_view(rowptr, colidx, self.nV_)
        q = dict(rp=np.asarray(rowptr, np.int32), ci=np.asarray(colidx, np.int32), order=np.arange(len(colidx), dtype=np.int32), q2c=q2c)
        if which == 0:
            self.qA, self.A = q, dict(p=cp, i=ri, order=q2c.astype(np.int32), zJ=len(colidx))
            self.Av = np.zeros((self.batch, len(colidx)))
        else:
            self.qH, self.H = q, dict(p=cp, i=ri, order=q2c.astype(np.int32), r=None, c=None, sym=False)
            self.Hv = np.zeros((self.batch, len(colidx)))
        self.set_csr_values(which, vals)

    def set_csr_values(self, which, vals):
        q = self.qA if which == 0 else self.qH
        v = self._bvals(vals, len(q["ci"]))
        (self.Av if which == 0 else self.Hv)[:, q["q2c"]] = v
        if self.first_solved:
            if which == 0:
                self.upd_A = True
            else:
                self.upd_H = True

    def get_csr(self, which):
        q = self.qA if which == 0 else self.qH
        vals = (self.Av if which == 0 else self.Hv)[:, q["q2c"]]
        return dict(RowIndex=np.asarray(q["rp"], np.int32), ColIndex=np.asarray(q["ci"], np.int32),
                    order=np.asarray(q["order"], np.int32), MatVal=vals.copy())

    def set_bounds_stacked(self, lb=None, ub=None):
        nV = self.nV_
        if lb is not None:
            lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), (self.batch, nV + self.nC_))
            self.lb[:], self.lbA[:] = lb[:, :nV], lb[:, nV:]
        if ub is not None:
            ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), (self.batch, nV + self.nC_))
            self.ub[:], self.ubA[:] = ub[:, :nV], ub[:, nV:]

    def get_bounds_stacked(self):
        return np.hstack([self.lb, self.lbA]), np.hstack([self.ub, self.ubA])

    def spmv(self, which, x, transpose=False):
        out = []
        for b in range(self.batch):
            if which == 0:
                out.append(orc.csc_times(self.nC_, self.nV_, self.A["p"], self.A["i"], self.Av[b], x[b], transpose=transpose))
            else:
                out.append(orc.csc_times(self.nV_, self.nV_, self.H["p"], self.H["i"], self.Hv[b], x[b]))
        return np.array(out).reshape(self.batch, -1)

    def vector_times(self, a, b):
        out = np.zeros(len(a))
        for i in range(np.shape(a)[1]):  # index order of Vector::times
            out = out + np.asarray(a)[:, i] * np.asarray(b)[:, i]
        return out

    def get_solution_stacked(self, want=("primal", "dual", "workingset")):
        Ax = self.spmv(0, self.x) if self.nC_ else np.zeros((self.batch, 0))
        return np.hstack([self.x, Ax]), self.y.copy(), -np.hstack([self.wb, self.wc]).astype(np.int32)

    # ---- solve: the init/hotstart decision of src/qpOASESInterface.cpp:141-211 per handle
    def _solve(self, active_mask):
        mode = "cold"
        if self.first_solved:
            varied = self.upd_A or self.upd_H
            if self.old_ms == MS_UNDEFINED:
                self.old_ms = MS_VARIED if varied else MS_FIXED
            else:
                if self.new_ms != MS_UNDEFINED:
                    self.old_ms = self.new_ms
                self.new_ms = MS_VARIED if varied else MS_FIXED
            if self.new_ms == MS_UNDEFINED:
                mode = "fixed" if self.old_ms == MS_FIXED else "varied"
            elif self.new_ms == MS_FIXED and self.old_ms == MS_FIXED:
                mode = "fixed"
            elif self.new_ms == MS_VARIED and self.old_ms == MS_VARIED:
                mode = "varied"
            else:
                mode = "reinit"  # status flip: init from the previous solution (:202-207)
                self.new_ms = self.old_ms = MS_UNDEFINED
        self.last_mode = mode
        for b in range(self.batch):
            if active_mask is not None and not active_mask[b]:
                continue
            s = self.solvers[b]
            Acsc = (self.A["p"], self.A["i"], self.Av[b])
            Hcsc = None if self.is_lp else (self.H["p"], self.H["i"], self.Hv[b])
            args = (self.g[b], self.lb[b], self.ub[b], self.lbA[b], self.ubA[b])
            its = 0
            if mode == "cold" or not self.inited[b]:
                st = s.init(Hcsc, args[0], Acsc, *args[1:], is_lp=self.is_lp)
                its = s.solution()[3]
            else:
                margs = (None if self.is_lp else self.Hv[b], self.Av[b]) + args
                st = s.hotstart(*args) if mode == "fixed" else (s.hotstart_matrices(*margs) if mode == "varied" else s.reinit(*margs))
                its = s.solution()[3]
            if st != 20:  # handle_error (src/qpOASESInterface.cpp:686-758), after an init as well as after a hot start
                st, added = s.handle_error()
                its += added
            self.inited[b] = (st == 20)
            x, y, obj, _ = s.solution()
            self.x[b], self.y[b], self.obj[b], self.status[b], self.iters[b] = x, y, obj, st, its
            wb, wc = s.working_set()
            self.wb[b], self.wc[b] = wb, wc
            Ax = orc.csc_times(self.nC_, self.nV_, *Acsc, x)
            Wb, Wc = orc.translate_working_set(wb, wc, x, Ax, *args[1:])
            _, self.kkt[b] = orc.kkt_residuals(self.nV_, self.nC_, Acsc, Hcsc, args[0], *args[1:], x, y, Wb, Wc)
        self.upd_A = self.upd_H = False
        self.first_solved = True

    def optimizeQP(self, stats=None, active_mask=None, maxiter=0):
        self._solve(active_mask)
        if stats is not None:
            stats.qp_iter_addValue(self.iters.copy())

    optimizeLP = optimizeQP

    # ---- results
    def get_optimal_solution(self): return self.x.copy()
    def get_obj_value(self): return self.obj.copy()
    def get_multipliers_bounds(self): return self.y[:, :self.nV_].copy()
    def get_multipliers_constr(self): return self.y[:, self.nV_:].copy()
    def get_status(self): return self.status.copy()
    def get_iterations(self): return self.iters.copy()

    def get_optimality_status(self, recompute=False):
        k = self.kkt
        return dict(primal_violation=k[:, 0], dual_violation=k[:, 1], stationarity_violation=k[:, 2], compl_violation=k[:, 3],
                    KKT_error=k[:, 4])

    def getLb(self): return self.lb.copy()
    def getUb(self): return self.ub.copy()
    def getLbA(self): return self.lbA.copy()
    def getUbA(self): return self.ubA.copy()
    def getG(self): return self.g.copy()
