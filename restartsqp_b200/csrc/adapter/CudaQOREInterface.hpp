// CudaQOREInterface.hpp -- the QORE-layout member of the drop-in plugin family: a
// QPSolverInterface backend with the data layout of the reference's QOREInterface (include/sqphot/QOREInterface.hpp:30-252) --
// row-compressed A and H, one lb / ub pair of length nVar_QP + nConstr_QP, x_qp = [x ; A x], one stacked multiplier vector --
// on the CUDA library (include/sqpb200.h: sqpb200_set_structure_*_csr, sqpb200_set_bounds_stacked,
// sqpb200_get_solution_stacked).  QPhandler's QORE branches (src/QPhandler.cpp:225-260, 369-383, 626-638) drive it unchanged:
// a maintainer adds `case CUDA_B200_QORE_LAYOUT:` next to `case QORE:` in the factory switch (src/QPhandler.cpp:58-76) and
// treats the enumerator like QORE in those three conditions.  One object = one QP instance.
// This file compiles against the reference's own headers.
#ifndef SQPHOTSTART_CUDAQOREINTERFACE_HPP
#define SQPHOTSTART_CUDAQOREINTERFACE_HPP

#include <sqphot/QPsolverInterface.hpp>
#include <sqpb200.h>

namespace SQPhotstart {

// include/sqphot/QOREInterface.hpp:16 declares INVALID_RETURN_TYPE for the same purpose; that header needs QORE's qpsolver.h, so
// this plugin brings its own exception class
DECLARE_STD_EXCEPTION(INVALID_RETURN_TYPE_QORE_LAYOUT);

class CudaQOREInterface : public QPSolverInterface {
public:
    /** Same construction convention as include/sqphot/QOREInterface.hpp:30-33. */
    CudaQOREInterface(NLPInfo nlp_info, QPType qptype, shared_ptr<const Options> options,
                      Ipopt::SmartPtr<Ipopt::Journalist> jnlst, int device = 0);
    /** Data constructor of the replay driver (include/sqphot/QOREInterface.hpp:35-41, src/QOREInterface.cpp:36-60):
     *  H and A row-compressed, lb and ub stacked. */
    CudaQOREInterface(shared_ptr<SpHbMat> H, shared_ptr<SpHbMat> A, shared_ptr<Vector> g, shared_ptr<Vector> lb,
                      shared_ptr<Vector> ub, shared_ptr<const Options> options = nullptr, int device = 0);
    ~CudaQOREInterface() override;

    void optimizeQP(shared_ptr<Stats> stats = nullptr) override;
    void optimizeLP(shared_ptr<Stats> stats = nullptr) override;

    double* get_optimal_solution() override { return x_qp_->values(); }                   // [x ; A x], callers read the first nVar_QP
    double get_obj_value() override { return obj_; }
    double* get_multipliers_bounds() override { return y_qp_->values(); }
    double* get_multipliers_constr() override { return y_qp_->values() + nVar_QP_; }
    void get_working_set(ActiveType* W_constr, ActiveType* W_bounds) override;
    /** QORE's own "workingset" vector of length nVar_QP + nConstr_QP (-1 upper, +1 lower, 0 inactive; src/QOREInterface.cpp:441). */
    const int* get_qore_working_set() const { return working_set_.data(); }
    Exitflag get_status() override;
    bool test_optimality(ActiveType* W_c = NULL, ActiveType* W_b = NULL) override;
    OptimalityStatus get_optimality_status() override { return qpOptimalStatus_; }

    // include/sqphot/QOREInterface.hpp:142-155: entries are clipped to +-INF; locations >= nVar_QP are constraint bounds
#ifndef SQPB200_QORE_NO_CLIP
    void set_g(int location, double value) override { g_->setValueAt(location, value < INF ? value : INF); g_dirty_ = true; }
    void set_lb(int location, double value) override { lb_->setValueAt(location, value > -INF ? value : -INF); b_dirty_ = true; }
    void set_ub(int location, double value) override { ub_->setValueAt(location, value < INF ? value : INF); b_dirty_ = true; }
#else  // test builds only (oracle/_ref/algorithm_nl_twin_noclip): bounds as the qpOASES-layout setters and the batched kernels take them
    void set_g(int location, double value) override { g_->setValueAt(location, value); g_dirty_ = true; }
    void set_lb(int location, double value) override { lb_->setValueAt(location, value); b_dirty_ = true; }
    void set_ub(int location, double value) override { ub_->setValueAt(location, value); b_dirty_ = true; }
#endif
    void set_g(shared_ptr<const Vector> rhs) override { g_->copy_vector(rhs->values()); g_dirty_ = true; }
    void set_lb(shared_ptr<const Vector> rhs) override { lb_->copy_vector(rhs->values()); b_dirty_ = true; }
    void set_ub(shared_ptr<const Vector> rhs) override { ub_->copy_vector(rhs->values()); b_dirty_ = true; }
    void set_lbA(int, double) override {}                    // :180-183: the constraint bounds live in lb / ub
    void set_lbA(shared_ptr<const Vector>) override {}
    void set_ubA(int, double) override {}
    void set_ubA(shared_ptr<const Vector>) override {}

    void set_H(shared_ptr<const SpTripletMat> rhs) override;
    void set_A(shared_ptr<const SpTripletMat> rhs, IdentityInfo I_info) override;
    void reset_constraints() override { lb_->set_zeros(); ub_->set_zeros(); b_dirty_ = true; }
    void WriteQPDataToFile(Ipopt::EJournalLevel level, Ipopt::EJournalCategory category, const string filename) override;

    const shared_ptr<Vector>& getLb() const override { return lb_; }
    const shared_ptr<Vector>& getUb() const override { return ub_; }
    const shared_ptr<Vector>& getLbA() const override { THROW_EXCEPTION(INVALID_RETURN_TYPE_QORE_LAYOUT, INVALID_RETURN_TYPE_MSG); }  // :124-130
    const shared_ptr<Vector>& getUbA() const override { THROW_EXCEPTION(INVALID_RETURN_TYPE_QORE_LAYOUT, INVALID_RETURN_TYPE_MSG); }
    const shared_ptr<Vector>& getG() const override { return g_; }
    shared_ptr<const SpHbMat> getH() const override { return H_; }
    shared_ptr<const SpHbMat> getA() const override { return A_; }

private:
    void create(int device);
    void solve(int mode, shared_ptr<Stats> stats);
    void check(int rc, const char* what);

    sqpb200_handle h_ = nullptr;
    int nVar_QP_ = 0, nConstr_QP_ = 0;
    QPType qptype_ = QP;
    shared_ptr<const Options> options_;
    Ipopt::SmartPtr<Ipopt::Journalist> jnlst_;
    shared_ptr<Vector> lb_, ub_, g_, x_qp_, y_qp_;  // lb_, ub_, x_qp_, y_qp_: nVar_QP + nConstr_QP entries (src/QOREInterface.cpp:199-204)
    shared_ptr<SpHbMat> A_, H_;                      // host mirrors of the row-compressed arrays for getA() / getH() / the dump
    std::vector<int> working_set_;
    bool g_dirty_ = true, b_dirty_ = true;
    bool A_structure_set_ = false, H_structure_set_ = false;
    double obj_ = 0.0;
    int status_ = QPERROR_NOTINITIALISED;
    OptimalityStatus qpOptimalStatus_;
};

}  // namespace SQPhotstart
#endif
