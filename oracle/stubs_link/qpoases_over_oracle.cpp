// A FUNCTIONAL stand-in for the part of qpOASES the reference calls (declarations: qpOASES.hpp of this directory), with the CPU
// oracle's active-set solver (oracle_qp.c) behind it.  TEST INFRASTRUCTURE ONLY.  Purpose: run the reference's REAL
// src/qpOASESInterface.cpp -- its init / hotstart state machine (:137-284, 817-833), handle_error (:686-758), get_working_set
// (:835-895), test_optimality (:498-684), get_status (:330-357) -- on a solver whose arithmetic is the oracle's, so that the
// restatements of those functions (oracle_l0.c, the backend state machines of oracle_sqp.c / capi_twin.cpp / the library's host
// side) can be compared with the reference's own code, call for call (oracle/backend_pin_test.cpp).  It is NOT qpOASES: the
// mapping of its entry points to the oracle is
//   init(H, g, A, lb, ub, lbA, ubA, nWSR)                         orc_qp_init (cold start)
//   init(..., nWSR, 0, x0)            (handle_error, infeasible)  orc_qp_handle_error(force_guess): init from [0; max(0,lbA); -min(0,ubA)]
//   init(..., nWSR, 0, xOpt, yOpt, &bounds)  (matrix status flip) orc_qp_reinit: init from the previous solution
//   hotstart(g, lb, ub, lbA, ubA, nWSR)                           orc_qp_hotstart
//   hotstart(H, g, A, lb, ub, lbA, ubA, nWSR)                     orc_qp_hotstart_matrices
#include <cstdio>
#include <cstdlib>
#include <qpOASES.hpp>
extern "C" {
#include "../oracle.h"
}

namespace qpOASES {

void Options::setToReliable() {}
Bounds::Bounds() {}
SparseMatrix::SparseMatrix(int_t nr, int_t nc, sparse_int_t* ir, sparse_int_t* jc, real_t* val) : nr_(nr), nc_(nc), ir_(ir), jc_(jc), val_(val) {}
SparseMatrix::~SparseMatrix() {}
sparse_int_t* SparseMatrix::createDiagInfo() { return 0; }
void SparseMatrix::setVal(const real_t* val) { val_ = val; }
returnValue SparseMatrix::print(const char*) const { return SUCCESSFUL_RETURN; }
SymSparseMat::SymSparseMat(int_t nr, int_t nc, sparse_int_t* ir, sparse_int_t* jc, real_t* val) : SparseMatrix(nr, nc, ir, jc, val) {}

SQProblem::SQProblem(int_t nV, int_t nC) : impl_(orc_qp_create(nV, nC)), nV_(nV), nC_(nC), status_(25), is_lp_(0) {}

static void finish(SQProblem* p, int st, int_t& nWSR, int iters) { p->status_ = st; nWSR = iters; }
static orc_qp_options opts(int_t nWSR) { orc_qp_options o; orc_qp_default_options(&o); o.max_iter = nWSR; return o; }
static int last_iters(const SQProblem* p) { int it = 0; orc_qp_get_solution((const orc_qp*)p->impl_, 0, 0, 0, &it); return it; }

returnValue SQProblem::init(SymSparseMat* H, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA,
                            const real_t* ubA, int_t& nWSR, real_t*, const real_t* xOpt, const real_t* yOpt, const Bounds* guessed) {
    orc_qp_options o = opts(nWSR);
    orc_qp* q = (orc_qp*)impl_;
    is_lp_ = H ? 0 : 1;
    int st, it;
    if (xOpt && yOpt && guessed) {  // init from the previous solution (src/qpOASESInterface.cpp:202-207)
        st = orc_qp_reinit(q, &o, H ? H->val_ : 0, A->val_, g, lb, ub, lbA, ubA);
        it = last_iters(this);
    } else if (xOpt) {              // init from the slack-feasible guess of handle_error (:716-729)
        int added = 0;
        st = orc_qp_handle_error(q, &o, 1, &added);
        it = added;
    } else {
        st = orc_qp_init(q, &o, H ? H->jc_ : 0, H ? H->ir_ : 0, H ? H->val_ : 0, g, A->jc_, A->ir_, A->val_, lb, ub, lbA, ubA, is_lp_);
        it = last_iters(this);
    }
    finish(this, st, nWSR, it);
    return st == 20 ? SUCCESSFUL_RETURN : RET_MAX_NWSR_REACHED;
}
returnValue SQProblem::init(int, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA, const real_t* ubA,
                            int_t& nWSR, real_t* t, const real_t* xOpt, const real_t* yOpt, const Bounds* guessed) {
    return init((SymSparseMat*)0, g, A, lb, ub, lbA, ubA, nWSR, t, xOpt, yOpt, guessed);
}
returnValue SQProblem::hotstart(const real_t* g, const real_t* lb, const real_t* ub, const real_t* lbA, const real_t* ubA, int_t& nWSR, real_t*) {
    orc_qp_options o = opts(nWSR);
    int st = orc_qp_hotstart((orc_qp*)impl_, &o, g, lb, ub, lbA, ubA);
    finish(this, st, nWSR, last_iters(this));
    return st == 20 ? SUCCESSFUL_RETURN : RET_MAX_NWSR_REACHED;
}
returnValue SQProblem::hotstart(SymSparseMat* H, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA,
                                const real_t* ubA, int_t& nWSR, real_t*) {
    orc_qp_options o = opts(nWSR);
    int st = orc_qp_hotstart_matrices((orc_qp*)impl_, &o, H ? H->val_ : 0, A->val_, g, lb, ub, lbA, ubA);
    finish(this, st, nWSR, last_iters(this));
    return st == 20 ? SUCCESSFUL_RETURN : RET_MAX_NWSR_REACHED;
}
returnValue SQProblem::hotstart(int, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA, const real_t* ubA,
                                int_t& nWSR, real_t* t) {
    return hotstart((SymSparseMat*)0, g, A, lb, ub, lbA, ubA, nWSR, t);
}
returnValue SQProblem::getPrimalSolution(real_t* x) const { orc_qp_get_solution((const orc_qp*)impl_, x, 0, 0, 0); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getDualSolution(real_t* y) const { orc_qp_get_solution((const orc_qp*)impl_, 0, y, 0, 0); return SUCCESSFUL_RETURN; }
real_t SQProblem::getObjVal() const {
    double obj = 0.0;
    orc_qp_get_solution((const orc_qp*)impl_, 0, 0, &obj, 0);
    return status_ == 20 ? obj : 1.0e20;  // INFTY for a problem that is not solved
}
QProblemStatus SQProblem::getStatus() const {
    switch (status_) {
    case 20: return QPS_SOLVED;
    case 25: return QPS_NOTINITIALISED;
    case 26: return QPS_PREPARINGAUXILIARYQP;
    case 27: return QPS_AUXILIARYQPSOLVED;
    case 29: return QPS_HOMOTOPYQPSOLVED;
    default: return QPS_PERFORMINGHOMOTOPY;  // 21 (internal error), 24 (iteration limit), 28: the state an interrupted homotopy leaves behind
    }
}
returnValue SQProblem::getBounds(Bounds&) const { return SUCCESSFUL_RETURN; }
returnValue SQProblem::getWorkingSetBounds(int_t* ws) const {
    int* wc = (int*)malloc(sizeof(int) * (nC_ > 0 ? nC_ : 1));
    orc_qp_get_working_set((const orc_qp*)impl_, ws, wc);
    free(wc);
    return SUCCESSFUL_RETURN;
}
returnValue SQProblem::getWorkingSetBounds(real_t* ws) const {
    int* wb = (int*)malloc(sizeof(int) * nV_); int* wc = (int*)malloc(sizeof(int) * (nC_ > 0 ? nC_ : 1));
    orc_qp_get_working_set((const orc_qp*)impl_, wb, wc);
    for (int i = 0; i < nV_; i++) ws[i] = wb[i];
    free(wb); free(wc);
    return SUCCESSFUL_RETURN;
}
returnValue SQProblem::getWorkingSetConstraints(int_t* ws) const {
    int* wb = (int*)malloc(sizeof(int) * nV_);
    orc_qp_get_working_set((const orc_qp*)impl_, wb, ws);
    free(wb);
    return SUCCESSFUL_RETURN;
}
returnValue SQProblem::getWorkingSetConstraints(real_t* ws) const {
    int* wb = (int*)malloc(sizeof(int) * nV_); int* wc = (int*)malloc(sizeof(int) * (nC_ > 0 ? nC_ : 1));
    orc_qp_get_working_set((const orc_qp*)impl_, wb, wc);
    for (int i = 0; i < nC_; i++) ws[i] = wc[i];
    free(wb); free(wc);
    return SUCCESSFUL_RETURN;
}
returnValue SQProblem::setOptions(const Options& o) { options_ = o; return SUCCESSFUL_RETURN; }
int_t SQProblem::getNV() const { return nV_; }
int_t SQProblem::getNC() const { return nC_; }
BooleanType SQProblem::isInfeasible() const { return status_ == 22 ? BT_TRUE : BT_FALSE; }
BooleanType SQProblem::isUnbounded() const { return status_ == 23 ? BT_TRUE : BT_FALSE; }
BooleanType SQProblem::isSolved() const { return status_ == 20 ? BT_TRUE : BT_FALSE; }

}  // namespace qpOASES
