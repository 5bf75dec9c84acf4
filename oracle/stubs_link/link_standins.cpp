// Definitions behind the stand-in headers qpOASES.hpp / qpsolver.h of this directory: every one aborts.  They exist so that the
// reference's own qpOASESInterface.cpp and QOREInterface.cpp LINK into oracle/_ref/qphandler_hs071 next to the CUDA plugins (the
// reference's QPhandler.cpp names their constructors in its factory switch); a run that ever reaches one of them is a test
// failure, not a solve.  TEST INFRASTRUCTURE ONLY.
#include <cstdio>
#include <cstdlib>
#include <qpOASES.hpp>
#include <qpsolver.h>

#define STANDIN(what)                                                                                             \
    do {                                                                                                          \
        fprintf(stderr, "link stand-in reached: %s (qpOASES / QORE are not in this container)\n", what);          \
        abort();                                                                                                  \
    } while (0)

#ifndef QPOASES_OVER_ORACLE  /* the functional stand-in (qpoases_over_oracle.cpp) defines these instead */
namespace qpOASES {
void Options::setToReliable() { STANDIN("qpOASES::Options::setToReliable"); }
Bounds::Bounds() {}
SparseMatrix::SparseMatrix(int_t, int_t, sparse_int_t*, sparse_int_t*, real_t*) { STANDIN("qpOASES::SparseMatrix"); }
SparseMatrix::~SparseMatrix() {}
sparse_int_t* SparseMatrix::createDiagInfo() { STANDIN("createDiagInfo"); return 0; }
void SparseMatrix::setVal(const real_t*) { STANDIN("setVal"); }
returnValue SparseMatrix::print(const char*) const { STANDIN("print"); return SUCCESSFUL_RETURN; }
SymSparseMat::SymSparseMat(int_t nr, int_t nc, sparse_int_t* ir, sparse_int_t* jc, real_t* val) : SparseMatrix(nr, nc, ir, jc, val) {}
SQProblem::SQProblem(int_t, int_t) { STANDIN("qpOASES::SQProblem"); }
returnValue SQProblem::init(SymSparseMat*, const real_t*, SparseMatrix*, const real_t*, const real_t*, const real_t*, const real_t*, int_t&, real_t*,
                            const real_t*, const real_t*, const Bounds*) { STANDIN("init"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::init(int, const real_t*, SparseMatrix*, const real_t*, const real_t*, const real_t*, const real_t*, int_t&, real_t*,
                            const real_t*, const real_t*, const Bounds*) { STANDIN("init"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::hotstart(const real_t*, const real_t*, const real_t*, const real_t*, const real_t*, int_t&, real_t*) { STANDIN("hotstart"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::hotstart(SymSparseMat*, const real_t*, SparseMatrix*, const real_t*, const real_t*, const real_t*, const real_t*, int_t&, real_t*) {
    STANDIN("hotstart"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::hotstart(int, const real_t*, SparseMatrix*, const real_t*, const real_t*, const real_t*, const real_t*, int_t&, real_t*) {
    STANDIN("hotstart"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getPrimalSolution(real_t*) const { STANDIN("getPrimalSolution"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getDualSolution(real_t*) const { STANDIN("getDualSolution"); return SUCCESSFUL_RETURN; }
real_t SQProblem::getObjVal() const { STANDIN("getObjVal"); return 0.0; }
QProblemStatus SQProblem::getStatus() const { STANDIN("getStatus"); return QPS_NOTINITIALISED; }
returnValue SQProblem::getBounds(Bounds&) const { STANDIN("getBounds"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getWorkingSetBounds(int_t*) const { STANDIN("getWorkingSetBounds"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getWorkingSetBounds(real_t*) const { STANDIN("getWorkingSetBounds"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getWorkingSetConstraints(int_t*) const { STANDIN("getWorkingSetConstraints"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::getWorkingSetConstraints(real_t*) const { STANDIN("getWorkingSetConstraints"); return SUCCESSFUL_RETURN; }
returnValue SQProblem::setOptions(const Options&) { STANDIN("setOptions"); return SUCCESSFUL_RETURN; }
int_t SQProblem::getNV() const { STANDIN("getNV"); return 0; }
int_t SQProblem::getNC() const { STANDIN("getNC"); return 0; }
BooleanType SQProblem::isInfeasible() const { STANDIN("isInfeasible"); return BT_FALSE; }
BooleanType SQProblem::isUnbounded() const { STANDIN("isUnbounded"); return BT_FALSE; }
BooleanType SQProblem::isSolved() const { STANDIN("isSolved"); return BT_FALSE; }
}  // namespace qpOASES
#endif

extern "C" {
qp_int QPNew(QoreProblem**, qp_int, qp_int, qp_int, qp_int) { STANDIN("QPNew"); return 0; }
void QPFree(QoreProblem**) { STANDIN("QPFree"); }
qp_int QPSetData(QoreProblem*, qp_int, qp_int, const qp_int*, const qp_int*, const double*, const qp_int*, const qp_int*, const double*) { STANDIN("QPSetData"); return 0; }
qp_int QPAdjust(QoreProblem*, double) { STANDIN("QPAdjust"); return 0; }
qp_int QPOptimize(QoreProblem*, const double*, const double*, const double*, const double*, const double*) { STANDIN("QPOptimize"); return 0; }
qp_int QPGetInt(QoreProblem*, const char*, qp_int*) { STANDIN("QPGetInt"); return 0; }
qp_int QPSetInt(QoreProblem*, const char*, qp_int) { STANDIN("QPSetInt"); return 0; }
qp_int QPGetDblVector(QoreProblem*, const char*, double*) { STANDIN("QPGetDblVector"); return 0; }
qp_int QPGetIntVector(QoreProblem*, const char*, qp_int*) { STANDIN("QPGetIntVector"); return 0; }
qp_int QPDataToFile(QoreProblem*, const char*) { STANDIN("QPDataToFile"); return 0; }
}
