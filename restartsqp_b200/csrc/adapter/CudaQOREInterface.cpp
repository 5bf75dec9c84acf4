// CudaQOREInterface.cpp -- see CudaQOREInterface.hpp.  Host glue only: the triplet -> compressed-row assembly, the value
// moves between the two storage orders, the active-set solve with its recovery path, the working-set translation and the
// KKT residuals all run in libsqpb200.so.
#include "CudaQOREInterface.hpp"

#include <cstdio>
#include <vector>

namespace SQPhotstart {

void CudaQOREInterface::check(int rc, const char* what) {
    if (rc < 0) {
        std::string msg = std::string(what) + ": " + (h_ ? sqpb200_last_error(h_) : "no handle");
        THROW_EXCEPTION(QP_INTERNAL_ERROR, msg);
    }
}

// src/QOREInterface.cpp:191-219 (allocate_memory) and :632-641 (set_solver_options: maxiter)
void CudaQOREInterface::create(int device) {
    sqpb200_options o;
    sqpb200_default_options(&o);
    if (options_) { o.qp_maxiter = options_->qp_maxiter; o.lp_maxiter = options_->lp_maxiter; }
    int rc = sqpb200_create(1, nVar_QP_, nConstr_QP_, qptype_ == LP ? SQPB200_LP : SQPB200_QP, device, &o, &h_);
    if (rc < 0) THROW_EXCEPTION(QP_INTERNAL_ERROR, "sqpb200_create failed: no CUDA device (there is no CPU fallback)");
    const int n = nVar_QP_ + nConstr_QP_;
    lb_ = make_shared<Vector>(n); ub_ = make_shared<Vector>(n); g_ = make_shared<Vector>(nVar_QP_);
    x_qp_ = make_shared<Vector>(n); y_qp_ = make_shared<Vector>(n);
    working_set_.assign(n, 0);
}

// src/QOREInterface.cpp:13-34
CudaQOREInterface::CudaQOREInterface(NLPInfo nlp_info, QPType qptype, shared_ptr<const Options> options,
                                     Ipopt::SmartPtr<Ipopt::Journalist> jnlst, int device)
    : qptype_(qptype), options_(options), jnlst_(jnlst) {
    nConstr_QP_ = nlp_info.nCon;
    nVar_QP_ = nlp_info.nVar + 2 * nlp_info.nCon;
    create(device);
}

// src/QOREInterface.cpp:36-60
CudaQOREInterface::CudaQOREInterface(shared_ptr<SpHbMat> H, shared_ptr<SpHbMat> A, shared_ptr<Vector> g, shared_ptr<Vector> lb,
                                     shared_ptr<Vector> ub, shared_ptr<const Options> options, int device)
    : qptype_(QP), options_(options) {
    nVar_QP_ = A->ColNum();
    nConstr_QP_ = A->RowNum();
    create(device);
    lb_->copy_vector(lb->values()); ub_->copy_vector(ub->values()); g_->copy_vector(g->values());
    A_ = A; H_ = H;
    // QPSetData(A_->RowIndex(), A_->ColIndex(), A_->MatVal(), H_->...) of :89-90: row pointers, column indices, values
    check(sqpb200_set_structure_csr(h_, SQPB200_MAT_A, A->EntryNum(), A->RowIndex(), A->ColIndex()), "set_structure_csr(A)");
    check(sqpb200_set_values_csr(h_, SQPB200_MAT_A, A->MatVal(), SQPB200_LOC_HOST, 0), "set_values_csr(A)");
    check(sqpb200_set_structure_csr(h_, SQPB200_MAT_H, H->EntryNum(), H->RowIndex(), H->ColIndex()), "set_structure_csr(H)");
    check(sqpb200_set_values_csr(h_, SQPB200_MAT_H, H->MatVal(), SQPB200_LOC_HOST, 0), "set_values_csr(H)");
    A_structure_set_ = H_structure_set_ = true;
}

CudaQOREInterface::~CudaQOREInterface() {
    if (h_) sqpb200_destroy(h_);
}

// src/QOREInterface.cpp:652-659
void CudaQOREInterface::set_H(shared_ptr<const SpTripletMat> rhs) {
    if (!H_structure_set_) {
        int nnz = sqpb200_set_structure_H_csr(h_, rhs->EntryNum(), rhs->RowIndex(), rhs->ColIndex(), rhs->isSymmetric() ? 1 : 0);
        check(nnz, "set_structure_H_csr");
        H_ = make_shared<SpHbMat>(nnz, nVar_QP_, nVar_QP_, true);
        check(sqpb200_get_structure_csr(h_, SQPB200_MAT_H, H_->RowIndex(), H_->ColIndex(), nullptr), "get_structure_csr(H)");
        H_structure_set_ = true;
    }
    check(sqpb200_set_values_H(h_, rhs->MatVal(), SQPB200_LOC_HOST, 0), "set_values_H");  // raises the matrix change flag (:656)
    check(sqpb200_get_values_csr(h_, SQPB200_MAT_H, H_->MatVal(), SQPB200_LOC_HOST), "get_values_csr(H)");
}

// src/QOREInterface.cpp:643-650
void CudaQOREInterface::set_A(shared_ptr<const SpTripletMat> rhs, IdentityInfo I_info) {
    if (!A_structure_set_) {
        int nnz = sqpb200_set_structure_A_csr(h_, rhs->EntryNum(), rhs->RowIndex(), rhs->ColIndex(), I_info.length, I_info.irow,
                                              I_info.jcol, I_info.size, I_info.value);
        check(nnz, "set_structure_A_csr");
        A_ = make_shared<SpHbMat>(nnz, nConstr_QP_, nVar_QP_, true);
        check(sqpb200_get_structure_csr(h_, SQPB200_MAT_A, A_->RowIndex(), A_->ColIndex(), nullptr), "get_structure_csr(A)");
        A_structure_set_ = true;
    }
    check(sqpb200_set_values_A(h_, rhs->MatVal(), SQPB200_LOC_HOST, 0), "set_values_A");
    check(sqpb200_get_values_csr(h_, SQPB200_MAT_A, A_->MatVal(), SQPB200_LOC_HOST), "get_values_csr(A)");
}

// QPSetData + QPOptimize(lb_, ub_, g_) + handle_error (src/QOREInterface.cpp:89-129, 607-629): the kernel's recovery path is the
// restart from the slack-feasible point that handle_error asks QORE for
void CudaQOREInterface::solve(int mode, shared_ptr<Stats> stats) {
    if (b_dirty_) { check(sqpb200_set_bounds_stacked(h_, lb_->values(), ub_->values(), SQPB200_LOC_HOST, 0), "set_bounds_stacked"); b_dirty_ = false; }
    if (g_dirty_) { check(sqpb200_set_vectors(h_, SQPB200_VEC_G, g_->values(), 0, nVar_QP_, SQPB200_LOC_HOST, 0), "set_vectors(g)"); g_dirty_ = false; }
    check(sqpb200_solve(h_, mode, 0, nullptr), "solve");
    int iters = 0;
    check(sqpb200_get_solution(h_, nullptr, nullptr, &obj_, &status_, &iters, SQPB200_LOC_HOST), "get_solution");
    check(sqpb200_get_solution_stacked(h_, x_qp_->values(), y_qp_->values(), working_set_.data(), SQPB200_LOC_HOST), "get_solution_stacked");
    if (stats != nullptr) stats->qp_iter_addValue(iters);  // :128-129
}

void CudaQOREInterface::optimizeQP(shared_ptr<Stats> stats) {
    solve(SQPB200_QP, stats);
    if (status_ != QP_OPTIMAL) THROW_EXCEPTION(QP_NOT_OPTIMAL, QP_NOT_OPTIMAL_MSG);  // :113-115
}

void CudaQOREInterface::optimizeLP(shared_ptr<Stats> stats) {
    solve(SQPB200_LP, stats);
    if (status_ != QP_OPTIMAL) THROW_EXCEPTION(LP_NOT_OPTIMAL, LP_NOT_OPTIMAL_MSG);  // :162-164
}

// src/QOREInterface.cpp:425-438
Exitflag CudaQOREInterface::get_status() {
    switch (status_) {
    case QP_OPTIMAL: case QPERROR_EXCEED_MAX_ITER: case QPERROR_INFEASIBLE: case QPERROR_UNBOUNDED: return (Exitflag)status_;
    default: return QPERROR_UNKNOWN;
    }
}

// src/QOREInterface.cpp:440-492: the translation of the stacked "workingset" vector (its constraint half with the comparison
// inside the fabs) gives entry for entry what the library's epilogue computes for the qpOASES twin (src/qpOASESInterface.cpp:835-895)
void CudaQOREInterface::get_working_set(ActiveType* W_constr, ActiveType* W_bounds) {
    std::vector<int> wb(nVar_QP_ > 0 ? nVar_QP_ : 1), wc(nConstr_QP_ > 0 ? nConstr_QP_ : 1);
    check(sqpb200_get_working_set(h_, wb.data(), wc.data(), 1, SQPB200_LOC_HOST), "get_working_set");
    for (int i = 0; i < nVar_QP_; i++) W_bounds[i] = (ActiveType)wb[i];
    for (int i = 0; i < nConstr_QP_; i++) W_constr[i] = (ActiveType)wc[i];
}

// src/QOREInterface.cpp:222-409: the four sums run over the stacked vectors, bounds first, x_qp_(nVar_QP + i) as the activity of
// constraint i: term for term the sums the solve kernel's epilogue evaluates; the tolerance is QORE's (:395)
bool CudaQOREInterface::test_optimality(ActiveType* W_c, ActiveType* W_b) {
    double out[5];
    check(sqpb200_kkt_residuals(h_, out, SQPB200_LOC_HOST), "kkt_residuals");
    if (W_c != NULL && W_b != NULL) get_working_set(W_c, W_b);
    qpOptimalStatus_.primal_violation = out[0];
    qpOptimalStatus_.dual_violation = out[1];
    qpOptimalStatus_.stationarity_violation = out[2];
    qpOptimalStatus_.compl_violation = out[3];
    qpOptimalStatus_.KKT_error = out[4];
    return !(out[4] > 1.0e-5);
}

// src/QOREInterface.cpp:582-598: sizes, lb, ub, g, then A and H as row pointers, column indices, values -- the `.log` layout the
// replay driver reads (test/QPsolvers_testers.cpp:48-150)
void CudaQOREInterface::WriteQPDataToFile(Ipopt::EJournalLevel level, Ipopt::EJournalCategory category, const string filename) {
    (void)level; (void)category;
    FILE* f = fopen(filename.c_str(), "w");
    if (!f) return;
    fprintf(f, "%d\n%d\n%d\n%d\n", nVar_QP_, nConstr_QP_, A_ ? A_->EntryNum() : 0, H_ ? H_->EntryNum() : 0);
    const shared_ptr<Vector>* vs[3] = {&lb_, &ub_, &g_};
    for (auto v : vs)
        for (int i = 0; i < (*v)->Dim(); i++) fprintf(f, "%23.16e\n", (*v)->values(i));
    const shared_ptr<SpHbMat>* ms[2] = {&A_, &H_};
    for (auto m : ms) {
        if (!*m) continue;
        for (int i = 0; i < (*m)->RowNum() + 1; i++) fprintf(f, "%d\n", (*m)->RowIndex(i));
        for (int i = 0; i < (*m)->EntryNum(); i++) fprintf(f, "%d\n", (*m)->ColIndex(i));
        for (int i = 0; i < (*m)->EntryNum(); i++) fprintf(f, "%23.16e\n", (*m)->MatVal(i));
    }
    fclose(f);
}

}  // namespace SQPhotstart
