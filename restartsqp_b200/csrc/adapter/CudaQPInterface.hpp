// CudaQPInterface.hpp -- the drop-in C++ plugin: a QPSolverInterface backend that sits beside
// qpOASESInterface / QOREInterface (include/sqphot/QPsolverInterface.hpp:43-194) and forwards to the C ABI of
// libsqpb200.so (include/sqpb200.h).  One object = one QP instance (batch of 1), which is what QPhandler
// constructs (src/QPhandler.cpp:58-76); the batched host driver talks to the C ABI directly.
//
// To register it a maintainer adds: an enumerator CUDA_B200 to `Solver`
// (include/sqphot/Types.hpp:91-97), a `case CUDA_B200:` to the factory switch (src/QPhandler.cpp:58-76), this
// header to include/sqphot/QPhandler.hpp:10-14, and `|| QPsolverChoice == CUDA_B200` to the condition at
// src/Algorithm.cpp:620.  This file compiles against the reference's own headers.
#ifndef SQPHOTSTART_CUDAQPINTERFACE_HPP
#define SQPHOTSTART_CUDAQPINTERFACE_HPP

#include <sqphot/QPsolverInterface.hpp>
#include <sqpb200.h>

namespace SQPhotstart {

class CudaQPInterface : public QPSolverInterface {
public:
    /** Same construction convention as include/sqphot/qpOASESInterface.hpp:39-41. */
    CudaQPInterface(NLPInfo nlp_info, QPType qptype, shared_ptr<const Options> options,
                    Ipopt::SmartPtr<Ipopt::Journalist> jnlst, int device = 0);
    /** Data constructor used by the replay driver (include/sqphot/qpOASESInterface.hpp:44-51). */
    CudaQPInterface(shared_ptr<SpHbMat> H, shared_ptr<SpHbMat> A, shared_ptr<Vector> g, shared_ptr<Vector> lb,
                    shared_ptr<Vector> ub, shared_ptr<Vector> lbA, shared_ptr<Vector> ubA,
                    shared_ptr<Options> options, int device = 0);
    ~CudaQPInterface() override;

    void optimizeQP(shared_ptr<Stats> stats) override;
    void optimizeLP(shared_ptr<Stats> stats) override;

    double* get_optimal_solution() override { return x_qp_->values(); }
    double get_obj_value() override { return obj_; }
    double* get_multipliers_bounds() override { return y_qp_->values(); }
    double* get_multipliers_constr() override { return y_qp_->values() + nVar_QP_; }
    void get_working_set(ActiveType* W_constr, ActiveType* W_bounds) override;
    Exitflag get_status() override { return (Exitflag)status_; }
    bool test_optimality(ActiveType* W_c = NULL, ActiveType* W_b = NULL) override;
    OptimalityStatus get_optimality_status() override { return qpOptimalStatus_; }

    void set_lb(int location, double value) override { lb_->setValueAt(location, value); dirty_[SQPB200_VEC_LB] = true; }
    void set_ub(int location, double value) override { ub_->setValueAt(location, value); dirty_[SQPB200_VEC_UB] = true; }
    void set_lbA(int location, double value) override { lbA_->setValueAt(location, value); dirty_[SQPB200_VEC_LBA] = true; }
    void set_ubA(int location, double value) override { ubA_->setValueAt(location, value); dirty_[SQPB200_VEC_UBA] = true; }
    void set_g(int location, double value) override { g_->setValueAt(location, value); dirty_[SQPB200_VEC_G] = true; }
    void set_ub(shared_ptr<const Vector> rhs) override { ub_->copy_vector(rhs->values()); dirty_[SQPB200_VEC_UB] = true; }
    void set_lb(shared_ptr<const Vector> rhs) override { lb_->copy_vector(rhs->values()); dirty_[SQPB200_VEC_LB] = true; }
    void set_lbA(shared_ptr<const Vector> rhs) override { lbA_->copy_vector(rhs->values()); dirty_[SQPB200_VEC_LBA] = true; }
    void set_ubA(shared_ptr<const Vector> rhs) override { ubA_->copy_vector(rhs->values()); dirty_[SQPB200_VEC_UBA] = true; }
    void set_g(shared_ptr<const Vector> rhs) override { g_->copy_vector(rhs->values()); dirty_[SQPB200_VEC_G] = true; }

    void set_H(shared_ptr<const SpTripletMat> rhs) override;
    void set_A(shared_ptr<const SpTripletMat> rhs, IdentityInfo I_info) override;
    void reset_constraints() override;
    void WriteQPDataToFile(Ipopt::EJournalLevel level, Ipopt::EJournalCategory category, const string filename) override;

    const shared_ptr<Vector>& getLb() const override { return lb_; }
    const shared_ptr<Vector>& getUb() const override { return ub_; }
    const shared_ptr<Vector>& getLbA() const override { return lbA_; }
    const shared_ptr<Vector>& getUbA() const override { return ubA_; }
    const shared_ptr<Vector>& getG() const override { return g_; }
    shared_ptr<const SpHbMat> getH() const override { return H_; }
    shared_ptr<const SpHbMat> getA() const override { return A_; }

private:
    void create(int device);
    void flush_vectors();
    void solve(int mode, shared_ptr<Stats> stats);
    void check(int rc, const char* what);

    sqpb200_handle h_ = nullptr;
    int nVar_QP_ = 0, nConstr_QP_ = 0;
    QPType qptype_ = QP;
    shared_ptr<const Options> options_;
    Ipopt::SmartPtr<Ipopt::Journalist> jnlst_;
    shared_ptr<Vector> lb_, ub_, lbA_, ubA_, g_, x_qp_, y_qp_;
    shared_ptr<SpHbMat> A_, H_;  // host mirrors of the CSC arrays for getA()/getH() and WriteQPDataToFile
    bool dirty_[5] = {true, true, true, true, true};
    bool A_structure_set_ = false, H_structure_set_ = false;
    double obj_ = 0.0;
    int status_ = QPERROR_NOTINITIALISED;
    OptimalityStatus qpOptimalStatus_;
};

}  // namespace SQPhotstart
#endif
