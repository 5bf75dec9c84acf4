// CudaQPInterface.cpp -- see CudaQPInterface.hpp.  Host glue only: every numerical step (triplet -> CSC
// assembly, value scatter, the active-set solve, working-set translation, KKT residuals) runs in libsqpb200.so.
#include "CudaQPInterface.hpp"

#include <vector>

namespace SQPhotstart {

void CudaQPInterface::check(int rc, const char* what) {
    if (rc < 0) {
        std::string msg = std::string(what) + ": " + (h_ ? sqpb200_last_error(h_) : "no handle");
        THROW_EXCEPTION(QP_INTERNAL_ERROR, msg);
    }
}

void CudaQPInterface::create(int device) {
    sqpb200_options o;
    sqpb200_default_options(&o);
    if (options_) { o.qp_maxiter = options_->qp_maxiter; o.lp_maxiter = options_->lp_maxiter; }
    int rc = sqpb200_create(1, nVar_QP_, nConstr_QP_, qptype_ == LP ? SQPB200_LP : SQPB200_QP, device, &o, &h_);
    if (rc < 0) THROW_EXCEPTION(QP_INTERNAL_ERROR, "sqpb200_create failed: no CUDA device (there is no CPU fallback)");
    lbA_ = make_shared<Vector>(nConstr_QP_); ubA_ = make_shared<Vector>(nConstr_QP_);
    lb_ = make_shared<Vector>(nVar_QP_); ub_ = make_shared<Vector>(nVar_QP_); g_ = make_shared<Vector>(nVar_QP_);
    x_qp_ = make_shared<Vector>(nVar_QP_); y_qp_ = make_shared<Vector>(nConstr_QP_ + nVar_QP_);
}

// src/qpOASESInterface.cpp:35-50, 106-128
CudaQPInterface::CudaQPInterface(NLPInfo nlp_info, QPType qptype, shared_ptr<const Options> options,
                                 Ipopt::SmartPtr<Ipopt::Journalist> jnlst, int device)
    : qptype_(qptype), options_(options), jnlst_(jnlst) {
    nConstr_QP_ = nlp_info.nCon;
    nVar_QP_ = nlp_info.nVar + 2 * nlp_info.nCon;
    create(device);
}

// src/qpOASESInterface.cpp:54-94
CudaQPInterface::CudaQPInterface(shared_ptr<SpHbMat> H, shared_ptr<SpHbMat> A, shared_ptr<Vector> g,
                                 shared_ptr<Vector> lb, shared_ptr<Vector> ub, shared_ptr<Vector> lbA,
                                 shared_ptr<Vector> ubA, shared_ptr<Options> options, int device)
    : qptype_(QP), options_(options) {
    nVar_QP_ = A->ColNum();
    nConstr_QP_ = A->RowNum();
    create(device);
    lb_->copy_vector(lb->values()); ub_->copy_vector(ub->values()); g_->copy_vector(g->values());
    lbA_->copy_vector(lbA->values()); ubA_->copy_vector(ubA->values());
    A_ = A; H_ = H;
    check(sqpb200_set_structure_csc(h_, SQPB200_MAT_A, A->EntryNum(), A->ColIndex(), A->RowIndex()), "set_structure_csc(A)");
    check(sqpb200_set_values_csc(h_, SQPB200_MAT_A, A->MatVal(), SQPB200_LOC_HOST, 0), "set_values_csc(A)");
    check(sqpb200_set_structure_csc(h_, SQPB200_MAT_H, H->EntryNum(), H->ColIndex(), H->RowIndex()), "set_structure_csc(H)");
    check(sqpb200_set_values_csc(h_, SQPB200_MAT_H, H->MatVal(), SQPB200_LOC_HOST, 0), "set_values_csc(H)");
    A_structure_set_ = H_structure_set_ = true;
}

CudaQPInterface::~CudaQPInterface() {
    if (h_) sqpb200_destroy(h_);
}

// src/qpOASESInterface.cpp:400-423
void CudaQPInterface::set_H(shared_ptr<const SpTripletMat> rhs) {
    if (!H_structure_set_) {
        int nnz = sqpb200_set_structure_H(h_, rhs->EntryNum(), rhs->RowIndex(), rhs->ColIndex(), rhs->isSymmetric() ? 1 : 0);
        check(nnz, "set_structure_H");
        H_ = make_shared<SpHbMat>(nnz, nVar_QP_, nVar_QP_, false);
        std::vector<int> order(nnz > 0 ? nnz : 1);
        check(sqpb200_get_structure(h_, SQPB200_MAT_H, H_->ColIndex(), H_->RowIndex(), order.data()), "get_structure(H)");
        H_structure_set_ = true;
    }
    check(sqpb200_set_values_H(h_, rhs->MatVal(), SQPB200_LOC_HOST, 0), "set_values_H");
    check(sqpb200_get_values_csc(h_, SQPB200_MAT_H, H_->MatVal(), SQPB200_LOC_HOST), "get_values_csc(H)");
}

// src/qpOASESInterface.cpp:426-442
void CudaQPInterface::set_A(shared_ptr<const SpTripletMat> rhs, IdentityInfo I_info) {
    if (!A_structure_set_) {
        int nnz = sqpb200_set_structure_A(h_, rhs->EntryNum(), rhs->RowIndex(), rhs->ColIndex(), I_info.length,
                                          I_info.irow, I_info.jcol, I_info.size, I_info.value);
        check(nnz, "set_structure_A");
        A_ = make_shared<SpHbMat>(nnz, nConstr_QP_, nVar_QP_, false);
        std::vector<int> order(nnz > 0 ? nnz : 1);
        check(sqpb200_get_structure(h_, SQPB200_MAT_A, A_->ColIndex(), A_->RowIndex(), order.data()), "get_structure(A)");
        A_structure_set_ = true;
    }
    check(sqpb200_set_values_A(h_, rhs->MatVal(), SQPB200_LOC_HOST, 0), "set_values_A");
    check(sqpb200_get_values_csc(h_, SQPB200_MAT_A, A_->MatVal(), SQPB200_LOC_HOST), "get_values_csc(A)");
}

void CudaQPInterface::flush_vectors() {
    struct { int which; shared_ptr<Vector>* v; int n; } vs[5] = {
        {SQPB200_VEC_G, &g_, nVar_QP_}, {SQPB200_VEC_LB, &lb_, nVar_QP_}, {SQPB200_VEC_UB, &ub_, nVar_QP_},
        {SQPB200_VEC_LBA, &lbA_, nConstr_QP_}, {SQPB200_VEC_UBA, &ubA_, nConstr_QP_}};
    for (auto& e : vs) {
        if (!dirty_[e.which] || e.n == 0) continue;
        // raises Update_g / Update_bounds inside the library exactly as the per-entry setters would
        check(sqpb200_set_vectors(h_, e.which, (*e.v)->values(), 0, e.n, SQPB200_LOC_HOST, 0), "set_vectors");
        dirty_[e.which] = false;
    }
}

void CudaQPInterface::solve(int mode, shared_ptr<Stats> stats) {
    flush_vectors();
    check(sqpb200_solve(h_, mode, 0, nullptr), "solve");
    int iters = 0;
    check(sqpb200_get_solution(h_, x_qp_->values(), y_qp_->values(), &obj_, &status_, &iters, SQPB200_LOC_HOST), "get_solution");
    if (stats != nullptr) stats->qp_iter_addValue(iters);  // src/qpOASESInterface.cpp:215-216
}

// src/qpOASESInterface.cpp:137-224
void CudaQPInterface::optimizeQP(shared_ptr<Stats> stats) {
    solve(SQPB200_QP, stats);
    if (status_ != QP_OPTIMAL) THROW_EXCEPTION(QP_NOT_OPTIMAL, QP_NOT_OPTIMAL_MSG);
}

// src/qpOASESInterface.cpp:227-284
void CudaQPInterface::optimizeLP(shared_ptr<Stats> stats) {
    solve(SQPB200_LP, stats);
    if (status_ != QP_OPTIMAL) THROW_EXCEPTION(LP_NOT_OPTIMAL, LP_NOT_OPTIMAL_MSG);
}

// src/qpOASESInterface.cpp:835-895
void CudaQPInterface::get_working_set(ActiveType* W_constr, ActiveType* W_bounds) {
    std::vector<int> wb(nVar_QP_ > 0 ? nVar_QP_ : 1), wc(nConstr_QP_ > 0 ? nConstr_QP_ : 1);
    check(sqpb200_get_working_set(h_, wb.data(), wc.data(), 1, SQPB200_LOC_HOST), "get_working_set");
    for (int i = 0; i < nVar_QP_; i++) W_bounds[i] = (ActiveType)wb[i];
    for (int i = 0; i < nConstr_QP_; i++) W_constr[i] = (ActiveType)wc[i];
}

// src/qpOASESInterface.cpp:498-684
bool CudaQPInterface::test_optimality(ActiveType* W_c, ActiveType* W_b) {
    double out[5];
    check(sqpb200_kkt_residuals(h_, out, SQPB200_LOC_HOST), "kkt_residuals");
    if (W_c != NULL && W_b != NULL) get_working_set(W_c, W_b);
    qpOptimalStatus_.primal_violation = out[0];
    qpOptimalStatus_.dual_violation = out[1];
    qpOptimalStatus_.stationarity_violation = out[2];
    qpOptimalStatus_.compl_violation = out[3];
    qpOptimalStatus_.KKT_error = out[4];
    return !(out[4] > 1.0e-6);
}

// src/qpOASESInterface.cpp:897-902
void CudaQPInterface::reset_constraints() {
    lb_->set_zeros(); ub_->set_zeros(); lbA_->set_zeros(); ubA_->set_zeros();
    dirty_[SQPB200_VEC_LB] = dirty_[SQPB200_VEC_UB] = dirty_[SQPB200_VEC_LBA] = dirty_[SQPB200_VEC_UBA] = true;
}

// src/qpOASESInterface.cpp:791-814
void CudaQPInterface::WriteQPDataToFile(Ipopt::EJournalLevel level, Ipopt::EJournalCategory category, const string filename) {
    (void)level; (void)category;
    FILE* f = fopen(("qpOASES" + filename).c_str(), "w");
    if (!f) return;
    const shared_ptr<Vector>* vs[5] = {&lb_, &lbA_, &ub_, &ubA_, &g_};
    for (auto v : vs)
        for (int i = 0; i < (*v)->Dim(); i++) fprintf(f, "%23.16e\n", (*v)->values(i));
    const shared_ptr<SpHbMat>* ms[2] = {&A_, &H_};
    for (auto m : ms) {
        if (!*m) continue;
        for (int i = 0; i < (*m)->EntryNum(); i++) fprintf(f, "%d\n", (*m)->RowIndex(i));
        for (int i = 0; i < (*m)->ColNum() + 1; i++) fprintf(f, "%d\n", (*m)->ColIndex(i));
        for (int i = 0; i < (*m)->EntryNum(); i++) fprintf(f, "%23.16e\n", (*m)->MatVal(i));
    }
    fclose(f);
}

}  // namespace SQPhotstart
