/* Stand-in for QORE's qpsolver.h (closed library, absent): the declarations src/QOREInterface.cpp uses, so that the reference's
 * own translation unit compiles and LINKS next to the CUDA plugins (oracle/Makefile target _ref/qphandler_hs071).  Every function
 * is defined in link_standins.cpp and aborts: nothing here solves anything.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_STUB_LINK_QPSOLVER_H
#define ORACLE_STUB_LINK_QPSOLVER_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int qp_int;
typedef struct QoreProblem_ QoreProblem;
enum { QPSOLVER_OK = 0, QPSOLVER_OPTIMAL = 1, QPSOLVER_ITER_LIMIT = 2, QPSOLVER_INFEASIBLE = 3, QPSOLVER_UNBOUNDED = 4 };
qp_int QPNew(QoreProblem** p, qp_int nvar, qp_int ncon, qp_int nnzA, qp_int nnzH);
void QPFree(QoreProblem** p);
qp_int QPSetData(QoreProblem* p, qp_int nvar, qp_int ncon, const qp_int* A_rp, const qp_int* A_ci, const double* A_val,
                 const qp_int* H_rp, const qp_int* H_ci, const double* H_val);
qp_int QPAdjust(QoreProblem* p, double flag);
qp_int QPOptimize(QoreProblem* p, const double* lb, const double* ub, const double* g, const double* x0, const double* y0);
qp_int QPGetInt(QoreProblem* p, const char* name, qp_int* value);
qp_int QPSetInt(QoreProblem* p, const char* name, qp_int value);
qp_int QPGetDblVector(QoreProblem* p, const char* name, double* v);
qp_int QPGetIntVector(QoreProblem* p, const char* name, qp_int* v);
qp_int QPDataToFile(QoreProblem* p, const char* filename);
#ifdef __cplusplus
}
#endif
#endif
