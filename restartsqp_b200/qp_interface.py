"""CudaQPInterface: the batched host-side mirror of the reference's QPSolverInterface plugin
(include/sqphot/QPsolverInterface.hpp:43-194) on top of the C ABI (include/sqpb200.h).

Method names, argument meaning and error behaviour follow the reference backend this one sits
beside, qpOASESInterface (src/qpOASESInterface.cpp); the only extension is the leading `batch`
dimension: one object owns `batch` independent QPs that share one sparsity pattern.  With
batch == 1 it behaves like one reference backend object, exceptions included.

All arithmetic happens in libsqpb200.so on the GPU.  There is no CPU path in this module.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from .sqp_types import (ActiveType, Exitflag, IdentityInfo, NLPInfo, Options, QPType, Stats, SpTripletMat,
                        QP_NOT_OPTIMAL, LP_NOT_OPTIMAL)


def _check(h, rc, what):
    if rc < 0:
        msg = capi.lib().sqpb200_last_error(h)
        raise capi.SqpB200Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
    return rc


class CudaQPInterface:
    """Batched QP/LP backend.  nV = nVar_QP, nC = nConstr_QP (src/qpOASESInterface.cpp:46-47)."""

    def __init__(self, nlp_info=None, qptype=QPType.QP, options=None, batch=1, device=0, nV=None, nC=None,
                 team_size=0, keep_state=True, factor_cap=0, debug_force_error_branch=False, refactorise_every=0):
        self.L = capi.lib()
        self.options = options if options is not None else Options()
        if nlp_info is not None:  # constructor (NLPInfo, QPType, options): src/qpOASESInterface.cpp:35-50
            nC = nlp_info.nCon
            nV = nlp_info.nVar + 2 * nlp_info.nCon
        self.nV_, self.nC_, self.batch, self.qptype, self.device = int(nV), int(nC), int(batch), QPType(qptype), int(device)
        o = capi.Options()
        self.L.sqpb200_default_options(C.byref(o))
        o.qp_maxiter, o.lp_maxiter = self.options.qp_maxiter, self.options.lp_maxiter
        o.team_size, o.keep_state, o.factor_cap = team_size, int(keep_state), int(factor_cap)
        o.debug_force_error_branch = int(bool(debug_force_error_branch))
        o.refactorise_every = int(refactorise_every)
        self.h = C.c_void_p()
        rc = self.L.sqpb200_create(self.batch, self.nV_, self.nC_, int(self.qptype), device, C.byref(o), C.byref(self.h))
        if rc < 0:
            raise capi.SqpB200Error("sqpb200_create failed (%d): no usable CUDA device or invalid sizes" % rc)
        self._A_set = self._H_set = False
        self._kkt = None

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.L.sqpb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(self.h, self.L.sqpb200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    # ------------------------------------------------------------------ setters (QPsolverInterface.hpp:144-165)
    def _set_vec(self, which, a0, a1=None):
        if a1 is None:  # vector form: [len] (broadcast) or [batch][len]
            v = capi.f64(a0)
            bc = int(v.ndim == 1)
            count = v.shape[-1]
            p, loc = capi.ptr(v)
            _check(self.h, self.L.sqpb200_set_vectors(self.h, which, p, 0, count, loc, bc), "set_vectors")
        else:  # (location, value) form: value scalar or [batch]
            v = np.ascontiguousarray(np.broadcast_to(np.asarray(a1, dtype=np.float64), (self.batch,)).reshape(self.batch, 1))
            p, loc = capi.ptr(v)
            _check(self.h, self.L.sqpb200_set_vectors(self.h, which, p, int(a0), 1, loc, 0), "set_vectors")

    def set_lb(self, a0, a1=None):
        self._set_vec(capi.VEC_LB, a0, a1)

    def set_ub(self, a0, a1=None):
        self._set_vec(capi.VEC_UB, a0, a1)

    def set_lbA(self, a0, a1=None):
        self._set_vec(capi.VEC_LBA, a0, a1)

    def set_ubA(self, a0, a1=None):
        self._set_vec(capi.VEC_UBA, a0, a1)

    def set_g(self, a0, a1=None):
        self._set_vec(capi.VEC_G, a0, a1)

    def set_H(self, rhs: SpTripletMat):
        """src/qpOASESInterface.cpp:400-423: first call builds the CSC structure, later calls refresh values."""
        if not self._H_set:
            r, c = capi.i32(rhs.RowIndex), capi.i32(rhs.ColIndex)
            _check(self.h, self.L.sqpb200_set_structure_H(self.h, len(r), r.ctypes.data_as(C.c_void_p),
                                                          c.ctypes.data_as(C.c_void_p), int(rhs.isSymmetric)),
                   "set_structure_H")
            self._H_set = True
        if rhs.MatVal is not None and rhs.EntryNum > 0:
            v = capi.f64(rhs.MatVal)
            p, loc = capi.ptr(v)
            _check(self.h, self.L.sqpb200_set_values_H(self.h, p, loc, int(v.ndim == 1)), "set_values_H")

    def set_A(self, rhs: SpTripletMat, I_info: IdentityInfo = None):
        """src/qpOASESInterface.cpp:426-442."""
        if not self._A_set:
            r, c = capi.i32(rhs.RowIndex), capi.i32(rhs.ColIndex)
            if I_info is None:
                I_info = IdentityInfo(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
            ir, jc, sz, val = capi.i32(I_info.irow), capi.i32(I_info.jcol), capi.i32(I_info.size), capi.f64(I_info.value)
            _check(self.h, self.L.sqpb200_set_structure_A(self.h, len(r), r.ctypes.data_as(C.c_void_p),
                                                          c.ctypes.data_as(C.c_void_p), len(sz),
                                                          ir.ctypes.data_as(C.c_void_p), jc.ctypes.data_as(C.c_void_p),
                                                          sz.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p)),
                   "set_structure_A")
            self._A_set = True
        if rhs.MatVal is not None:
            v = capi.f64(rhs.MatVal) if rhs.EntryNum > 0 else np.zeros(1)  # an empty Jacobian still raises Update_A
            p, loc = capi.ptr(v)
            _check(self.h, self.L.sqpb200_set_values_A(self.h, p, loc, int(v.ndim == 1)), "set_values_A")

    def set_csc(self, which, colptr, rowidx, vals):
        """Data-constructor path (src/qpOASESInterface.cpp:54-94): CSC arrays given directly."""
        cp, ri = capi.i32(colptr), capi.i32(rowidx)
        _check(self.h, self.L.sqpb200_set_structure_csc(self.h, which, len(ri), cp.ctypes.data_as(C.c_void_p),
                                                        ri.ctypes.data_as(C.c_void_p)), "set_structure_csc")
        if which == capi.MAT_A:
            self._A_set = True
        else:
            self._H_set = True
        self.set_csc_values(which, vals)

    def set_csc_values(self, which, vals):
        v = capi.f64(vals)
        if v.shape[-1] == 0:  # empty matrix: the call only raises the Update_A / Update_H flag, like the reference's setters
            v = np.zeros(1)
        p, loc = capi.ptr(v)
        _check(self.h, self.L.sqpb200_set_values_csc(self.h, which, p, loc, int(v.ndim == 1)), "set_values_csc")

    # ------------------------------------------------------------------ QORE layout (compressed-row matrices, stacked vectors)
    def set_A_csr(self, rhs: SpTripletMat, I_info: IdentityInfo = None):
        """QOREInterface::set_A (src/QOREInterface.cpp:643-650): the first call builds the compressed-row structure
        (SpHbMat(..., isCompressedRow = true)::setStructure, src/SpHbMat.cpp:238-250), later calls refresh values."""
        if not self._A_set:
            r, c = capi.i32(rhs.RowIndex), capi.i32(rhs.ColIndex)
            if I_info is None:
                I_info = IdentityInfo(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
            ir, jc, sz, val = capi.i32(I_info.irow), capi.i32(I_info.jcol), capi.i32(I_info.size), capi.f64(I_info.value)
            _check(self.h, self.L.sqpb200_set_structure_A_csr(self.h, len(r), r.ctypes.data_as(C.c_void_p),
                                                              c.ctypes.data_as(C.c_void_p), len(sz),
                                                              ir.ctypes.data_as(C.c_void_p), jc.ctypes.data_as(C.c_void_p),
                                                              sz.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p)),
                   "set_structure_A_csr")
            self._A_set = True
        self.set_A(rhs, I_info)  # values: triplet order, scattered on the device

    def set_H_csr(self, rhs: SpTripletMat):
        """QOREInterface::set_H (src/QOREInterface.cpp:652-659)."""
        if not self._H_set:
            r, c = capi.i32(rhs.RowIndex), capi.i32(rhs.ColIndex)
            _check(self.h, self.L.sqpb200_set_structure_H_csr(self.h, len(r), r.ctypes.data_as(C.c_void_p),
                                                              c.ctypes.data_as(C.c_void_p), int(rhs.isSymmetric)),
                   "set_structure_H_csr")
            self._H_set = True
        self.set_H(rhs)

    def set_csr(self, which, rowptr, colidx, vals):
        """Data-constructor path of the QORE backend (src/QOREInterface.cpp:36-60): compressed-row arrays given directly."""
        rp, ci = capi.i32(rowptr), capi.i32(colidx)
        _check(self.h, self.L.sqpb200_set_structure_csr(self.h, which, len(ci), rp.ctypes.data_as(C.c_void_p),
                                                        ci.ctypes.data_as(C.c_void_p)), "set_structure_csr")
        if which == capi.MAT_A:
            self._A_set = True
        else:
            self._H_set = True
        self.set_csr_values(which, vals)

    def set_csr_values(self, which, vals):
        v = capi.f64(vals)
        if v.shape[-1] == 0:
            v = np.zeros(1)
        p, loc = capi.ptr(v)
        _check(self.h, self.L.sqpb200_set_values_csr(self.h, which, p, loc, int(v.ndim == 1)), "set_values_csr")

    def get_csr(self, which):
        """Compressed-row arrays as QPSetData receives them (src/QOREInterface.cpp:89-90): RowIndex = row pointers,
        ColIndex = column of each entry, MatVal[batch][nnz] in that storage order, order = triplet entry -> position."""
        nnz = self.L.sqpb200_get_nnz(self.h, which)
        nrow = self.nC_ if which == capi.MAT_A else self.nV_
        rp, ci, od = np.zeros(nrow + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.int32)
        _check(self.h, self.L.sqpb200_get_structure_csr(self.h, which, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                                        od.ctypes.data_as(C.c_void_p)), "get_structure_csr")
        vals = np.zeros((self.batch, nnz))
        if nnz:
            _check(self.h, self.L.sqpb200_get_values_csr(self.h, which, vals.ctypes.data_as(C.c_void_p), capi.LOC_HOST), "get_values_csr")
        return dict(RowIndex=rp, ColIndex=ci, order=od, MatVal=vals)

    def set_bounds_stacked(self, lb=None, ub=None):
        """lb / ub [nV+nC] (shared) or [batch][nV+nC]: variable bounds first, then constraint bounds."""
        arrs = [None if a is None else capi.f64(a) for a in (lb, ub)]
        given = [a for a in arrs if a is not None]
        if not given:
            return
        assert all(a.ndim == given[0].ndim and a.shape[-1] == self.nV_ + self.nC_ for a in given)
        (pl, l1), (pu, l2) = capi.ptr(arrs[0]), capi.ptr(arrs[1])
        _check(self.h, self.L.sqpb200_set_bounds_stacked(self.h, pl, pu, l1 if arrs[0] is not None else l2, int(given[0].ndim == 1)),
               "set_bounds_stacked")

    def get_bounds_stacked(self):
        lb, ub = np.empty((self.batch, self.nV_ + self.nC_)), np.empty((self.batch, self.nV_ + self.nC_))
        _check(self.h, self.L.sqpb200_get_bounds_stacked(self.h, lb.ctypes.data_as(C.c_void_p), ub.ctypes.data_as(C.c_void_p),
                                                         capi.LOC_HOST), "get_bounds_stacked")
        return lb, ub

    def get_solution_stacked(self, want=("primal", "dual", "workingset")):
        """(primal [x ; A x], dual [bound ; constraint multipliers], workingset in QORE's sign) as [batch][nV+nC] arrays."""
        n = self.nV_ + self.nC_
        pr = np.empty((self.batch, n)) if "primal" in want else None
        du = np.empty((self.batch, n)) if "dual" in want else None
        ws = np.empty((self.batch, n), np.int32) if "workingset" in want else None
        pp = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in (pr, du, ws)]
        _check(self.h, self.L.sqpb200_get_solution_stacked(self.h, *pp, capi.LOC_HOST), "get_solution_stacked")
        return pr, du, ws

    # ------------------------------------------------------------------ batched QPhandler data kernels
    def qphandler_bounds(self, mode, n, m, delta, x_l, x_u, x_k, c_l=None, c_u=None, c_k=None):
        """QPhandler::set_bounds (mode 0) / update_bounds (1) / update_delta (2) on the device, per-instance delta
        (src/QPhandler.cpp:167-261, 342-419, 533-567).  Arrays are [batch][n] / [batch][m] (numpy or CUDA tensors)."""
        args = [delta, x_l, x_u, x_k] + ([c_l, c_u, c_k] if (mode != 2 and m > 0) else [None, None, None])
        ptrs, loc = [], capi.LOC_HOST
        for a in args:
            p, l = capi.ptr(a)
            ptrs.append(p)
            if a is not None:
                loc = l
        _check(self.h, self.L.sqpb200_qphandler_bounds(self.h, mode, n, m, *ptrs, loc), "qphandler_bounds")

    def qphandler_g(self, n, m, grad, rho):
        """g = [grad ; rho*1] (src/QPhandler.cpp:272-297, 430-463); grad [batch][n] or None, rho [batch] or None."""
        pg, l1 = capi.ptr(grad)
        pr, l2 = capi.ptr(rho)
        _check(self.h, self.L.sqpb200_qphandler_g(self.h, n, m, pg, pr, l1 if grad is not None else l2), "qphandler_g")

    def reset_constraints(self):
        """src/qpOASESInterface.cpp:897-902."""
        for w, n in ((capi.VEC_LB, self.nV_), (capi.VEC_UB, self.nV_), (capi.VEC_LBA, self.nC_), (capi.VEC_UBA, self.nC_)):
            if n:
                self._set_vec(w, np.zeros(n))

    # ------------------------------------------------------------------ solve (QPsolverInterface.hpp:74-85)
    def _solve(self, mode, stats, active_mask, maxiter):
        m = None
        if active_mask is not None:
            m = np.ascontiguousarray(active_mask, dtype=np.uint8)
        _check(self.h, self.L.sqpb200_solve(self.h, int(mode), int(maxiter), None if m is None else m.ctypes.data_as(C.c_void_p)),
               "solve")
        self._kkt = None
        if stats is not None:
            stats.qp_iter_addValue(self.get_iterations())

    def optimizeQP(self, stats: Stats = None, active_mask=None, maxiter=0):
        """src/qpOASESInterface.cpp:137-224.  With batch == 1 raises QP_NOT_OPTIMAL like the reference."""
        self._solve(QPType.QP, stats, active_mask, maxiter)
        if self.batch == 1 and int(self.get_status()[0]) != Exitflag.QP_OPTIMAL:
            raise QP_NOT_OPTIMAL("QP solver reports status %d" % int(self.get_status()[0]))

    def optimizeLP(self, stats: Stats = None, active_mask=None, maxiter=0):
        """src/qpOASESInterface.cpp:227-284."""
        self._solve(QPType.LP, stats, active_mask, maxiter)
        if self.batch == 1 and int(self.get_status()[0]) != Exitflag.QP_OPTIMAL:
            raise LP_NOT_OPTIMAL("LP solver reports status %d" % int(self.get_status()[0]))

    def io_layout(self):
        """Byte layout of the handle's input block (g, lb, ub, lbA, ubA) and result block (x, y, obj, kkt, status, iters)."""
        ino, outo = (C.c_size_t * 5)(), (C.c_size_t * 6)()
        inb, outb = C.c_size_t(), C.c_size_t()
        _check(self.h, self.L.sqpb200_io_layout(self.h, ino, C.byref(inb), outo, C.byref(outb)), "io_layout")
        return dict(in_off=dict(zip(("g", "lb", "ub", "lbA", "ubA"), list(ino))), in_bytes=inb.value,
                    out_off=dict(zip(("x", "y", "obj", "kkt", "status", "iters"), list(outo))), out_bytes=outb.value)

    def solve_host(self, mode, in_block=None, Aval=None, Hval=None, out_block=None, maxiter=0):
        """One call per solve for host-resident (pinned) data: uploads, solve and the result download are queued on the
        handle's stream; nothing is waited for.  Arguments are raw host addresses (ints) or None."""
        _check(self.h, self.L.sqpb200_solve_host(self.h, int(mode), int(maxiter), in_block, Aval, Hval, out_block), "solve_host")
        self._kkt = None

    def synchronize(self):
        _check(self.h, self.L.sqpb200_synchronize(self.h), "synchronize")

    # ------------------------------------------------------------------ getters (QPsolverInterface.hpp:97-135)
    def _get_solution(self, want):
        B, nV, nC = self.batch, self.nV_, self.nC_
        x = np.empty((B, nV)) if "x" in want else None
        y = np.empty((B, nV + nC)) if "y" in want else None
        obj = np.empty(B) if "obj" in want else None
        st = np.empty(B, np.int32) if "status" in want else None
        it = np.empty(B, np.int32) if "iters" in want else None
        pp = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in (x, y, obj, st, it)]
        _check(self.h, self.L.sqpb200_get_solution(self.h, *pp, capi.LOC_HOST), "get_solution")
        return x, y, obj, st, it

    def get_optimal_solution(self):
        return self._get_solution("x")[0]

    def get_obj_value(self):
        return self._get_solution(("obj",))[2]

    def get_multipliers_bounds(self):  # y[0:nV]
        return self._get_solution("y")[1][:, :self.nV_]

    def get_multipliers_constr(self):  # y + nVar_QP_, src/qpOASESInterface.cpp:303-305
        return self._get_solution("y")[1][:, self.nV_:]

    def get_status(self):
        return self._get_solution(("status",))[3]

    def get_iterations(self):
        return self._get_solution(("iters",))[4]

    def get_working_set(self, translated=True):
        """(W_constr, W_bounds) as ActiveType values (src/qpOASESInterface.cpp:835-895); translated=False
        returns the raw qpOASES convention (+1 upper, -1 lower, 0 inactive)."""
        wb = np.empty((self.batch, self.nV_), np.int32)
        wc = np.empty((self.batch, self.nC_), np.int32)
        _check(self.h, self.L.sqpb200_get_working_set(self.h, wb.ctypes.data_as(C.c_void_p), wc.ctypes.data_as(C.c_void_p),
                                                      int(translated), capi.LOC_HOST), "get_working_set")
        return wc, wb

    def get_optimality_status(self, recompute=False):
        """OptimalityStatus fields as arrays [batch] (include/sqphot/Types.hpp:107-119)."""
        out = np.empty((self.batch, 5))
        f = self.L.sqpb200_kkt_residuals_recompute if recompute else self.L.sqpb200_kkt_residuals
        _check(self.h, f(self.h, out.ctypes.data_as(C.c_void_p), capi.LOC_HOST), "kkt_residuals")
        self._kkt = out
        return dict(primal_violation=out[:, 0], dual_violation=out[:, 1], stationarity_violation=out[:, 2],
                    compl_violation=out[:, 3], KKT_error=out[:, 4])

    def test_optimality(self, recompute=False):
        """KKT_error <= 1e-6 per instance (src/qpOASESInterface.cpp:673)."""
        return self.get_optimality_status(recompute)["KKT_error"] <= 1.0e-6

    def _get_vec(self, which, n):
        out = np.empty((self.batch, n))
        if n:
            _check(self.h, self.L.sqpb200_get_vectors(self.h, which, out.ctypes.data_as(C.c_void_p), capi.LOC_HOST), "get_vectors")
        return out

    def getLb(self):
        return self._get_vec(capi.VEC_LB, self.nV_)

    def getUb(self):
        return self._get_vec(capi.VEC_UB, self.nV_)

    def getLbA(self):
        return self._get_vec(capi.VEC_LBA, self.nC_)

    def getUbA(self):
        return self._get_vec(capi.VEC_UBA, self.nC_)

    def getG(self):
        return self._get_vec(capi.VEC_G, self.nV_)

    def _get_mat(self, which):
        nnz = self.L.sqpb200_get_nnz(self.h, which)
        cp, ri, od = np.zeros(self.nV_ + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.int32)
        _check(self.h, self.L.sqpb200_get_structure(self.h, which, cp.ctypes.data_as(C.c_void_p), ri.ctypes.data_as(C.c_void_p),
                                                    od.ctypes.data_as(C.c_void_p)), "get_structure")
        vals = np.zeros((self.batch, nnz))
        if nnz:
            _check(self.h, self.L.sqpb200_get_values_csc(self.h, which, vals.ctypes.data_as(C.c_void_p), capi.LOC_HOST), "get_values_csc")
        return dict(ColIndex=cp, RowIndex=ri, order=od, MatVal=vals)

    def getA(self):
        """CSC arrays of A (the SpHbMat the reference hands to qpOASES, src/qpOASESInterface.cpp:432-436)."""
        return self._get_mat(capi.MAT_A)

    def getH(self):
        return self._get_mat(capi.MAT_H)

    def spmv(self, which, x, transpose=False):
        """Batched SpHbMat::times / transposed_times on the handle's matrices."""
        x = capi.f64(x)
        nout = (self.nV_ if transpose else self.nC_) if which == capi.MAT_A else self.nV_
        if capi._is_torch(x):
            import torch
            y = torch.empty((self.batch, nout), dtype=torch.float64, device=x.device)
        else:
            y = np.empty((self.batch, nout))
        px, loc = capi.ptr(x)
        py, _ = capi.ptr(y)
        _check(self.h, self.L.sqpb200_spmv(self.h, which, int(transpose), px, py, loc), "spmv")
        return y

    def vector_times(self, a, b):
        """Batched Vector::times (src/Vector.cpp:237-251): out[batch] = sum_i a[b][i] * b[b][i] in index order, on the device."""
        a, b = capi.f64(a), capi.f64(b)
        out = np.empty(a.shape[0])
        (pa, loc), (pb, _) = capi.ptr(a), capi.ptr(b)
        rc = self.L.sqpb200_vector_reduce(self.device, 2, a.shape[0], a.shape[1], pa, pb, out.ctypes.data_as(C.c_void_p), loc, None)
        if rc < 0:
            raise capi.SqpB200Error("vector_reduce failed (%d)" % rc)
        return out

    def launch_count(self):
        return int(self.L.sqpb200_launch_count(self.h))

    def last_solve_ms(self):
        return float(self.L.sqpb200_last_solve_ms(self.h))

    PROFILE_PHASES = ("stepdir", "ratio", "step", "remove", "extend_R", "w_vec", "add", "refac_W", "refac_M", "refac_chol",
                      "drift", "ensure_li", "R_solves(in stepdir)", "setup", "epilogue", "total")

    def profile(self, reset=True):
        """Per-phase cycle counters of a -DQP_PROFILE build (all zero otherwise), summed over instances."""
        out = np.zeros(16, np.int64)
        _check(self.h, self.L.sqpb200_get_profile(self.h, out.ctypes.data_as(C.c_void_p), int(reset)), "get_profile")
        return dict(zip(self.PROFILE_PHASES, out.tolist()))

    def solve_config(self):
        t, q, s = C.c_int(), C.c_int(), C.c_int()
        _check(self.h, self.L.sqpb200_solve_config(self.h, C.byref(t), C.byref(q), C.byref(s)), "solve_config")
        return dict(team_size=t.value, qps_per_cta=q.value, smem_per_cta=s.value)

    def WriteQPDataToFile(self, filename, instance=0):
        """qpOASES-layout dump of one instance (src/qpOASESInterface.cpp:791-814 with the QPOASES branches of
        Vector::write_to_file / SpHbMat::write_to_file, src/SpHbMat.cpp:568-578): lb, lbA, ub, ubA, g, A, H."""
        A, H = self.getA(), (self.getH() if self._H_set else None)
        with open("qpOASES" + filename, "w") as f:
            for v in (self.getLb(), self.getLbA(), self.getUb(), self.getUbA(), self.getG()):
                for t in v[instance]:
                    f.write("%23.16e\n" % t)
            for M in (A, H):
                if M is None:
                    continue
                for t in M["RowIndex"]:
                    f.write("%d\n" % t)
                for t in M["ColIndex"]:
                    f.write("%d\n" % t)
                for t in M["MatVal"][instance]:
                    f.write("%23.16e\n" % t)
