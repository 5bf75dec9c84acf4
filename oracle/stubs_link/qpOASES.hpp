// Stand-in for qpOASES 3.2.1's qpOASES.hpp (fetched at configure time by the reference, absent here): the declarations
// src/qpOASESInterface.cpp uses, so that the reference's own translation unit compiles and LINKS next to the CUDA plugins
// (oracle/Makefile target _ref/qphandler_hs071).  Every member is defined in link_standins.cpp and aborts: nothing here solves
// anything, and nothing in it is taken from qpOASES beyond the names the reference calls.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_LINK_QPOASES_HPP
#define ORACLE_STUB_LINK_QPOASES_HPP
#define ORACLE_STUB_QPOASES_HPP  // keeps oracle/stubs/qpOASES.hpp out
namespace qpOASES {
typedef double real_t;
typedef int int_t;
typedef int sparse_int_t;
enum BooleanType { BT_FALSE = 0, BT_TRUE = 1 };
enum PrintLevel { PL_DEBUG_ITER = -2, PL_TABULAR, PL_NONE, PL_LOW, PL_MEDIUM, PL_HIGH };
enum QProblemStatus { QPS_NOTINITIALISED, QPS_PREPARINGAUXILIARYQP, QPS_AUXILIARYQPSOLVED, QPS_PERFORMINGHOMOTOPY, QPS_HOMOTOPYQPSOLVED, QPS_SOLVED };
enum returnValue { SUCCESSFUL_RETURN = 0, RET_MAX_NWSR_REACHED = 64 };
struct Options {
    PrintLevel printLevel;
    void setToReliable();
};
class Bounds {
public:
    Bounds();
};
class SparseMatrix {
public:
    SparseMatrix(int_t nr, int_t nc, sparse_int_t* ir, sparse_int_t* jc, real_t* val);
    virtual ~SparseMatrix();
    sparse_int_t* createDiagInfo();
    void setVal(const real_t* val);
    returnValue print(const char* name = 0) const;
    // used by the functional stand-in only (qpoases_over_oracle.cpp): the column-compressed arrays the reference hands over
    int_t nr_, nc_;
    sparse_int_t *ir_, *jc_;
    const real_t* val_;
};
class SymSparseMat : public SparseMatrix {
public:
    SymSparseMat(int_t nr, int_t nc, sparse_int_t* ir, sparse_int_t* jc, real_t* val);
};
class SQProblem {
public:
    SQProblem(int_t nV, int_t nC);
    returnValue init(SymSparseMat* H, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA,
                     const real_t* ubA, int_t& nWSR, real_t* cputime = 0, const real_t* xOpt = 0, const real_t* yOpt = 0,
                     const Bounds* guessedBounds = 0);
    returnValue init(int H, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA,
                     const real_t* ubA, int_t& nWSR, real_t* cputime = 0, const real_t* xOpt = 0, const real_t* yOpt = 0,
                     const Bounds* guessedBounds = 0);
    returnValue hotstart(const real_t* g, const real_t* lb, const real_t* ub, const real_t* lbA, const real_t* ubA, int_t& nWSR,
                         real_t* cputime = 0);
    returnValue hotstart(SymSparseMat* H, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA,
                         const real_t* ubA, int_t& nWSR, real_t* cputime = 0);
    returnValue hotstart(int H, const real_t* g, SparseMatrix* A, const real_t* lb, const real_t* ub, const real_t* lbA,
                         const real_t* ubA, int_t& nWSR, real_t* cputime = 0);
    returnValue getPrimalSolution(real_t* x) const;
    returnValue getDualSolution(real_t* y) const;
    real_t getObjVal() const;
    QProblemStatus getStatus() const;
    returnValue getBounds(Bounds& b) const;
    returnValue getWorkingSetBounds(int_t* ws) const;
    returnValue getWorkingSetBounds(real_t* ws) const;
    returnValue getWorkingSetConstraints(int_t* ws) const;
    returnValue getWorkingSetConstraints(real_t* ws) const;
    returnValue setOptions(const Options& o);
    int_t getNV() const;
    int_t getNC() const;
    BooleanType isInfeasible() const;
    BooleanType isUnbounded() const;
    BooleanType isSolved() const;
    // used by the functional stand-in only
    void* impl_;
    int_t nV_, nC_;
    int status_, is_lp_;
    Options options_;
};
}  // namespace qpOASES
#endif
