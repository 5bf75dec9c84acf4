/* oracle/oracle.h -- CPU ORACLE: TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the RestartSQP hot path, used as the
 * checker for the CUDA kernels.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (restartsqp_b200/lib/libsqpb200.so) never links, loads or calls it.
 *
 * Pinning status:
 *   - L0 functions (triplet->CSC, value scatter, SpMV/SpMTV, norms): PINNED against the
 *     reference's own code compiled in oracle/_ref (tests/test_oracle_l0.py) and against
 *     committed golden vectors generated from it (tests/golden/l0_golden.json).
 *   - L1/L2 glue (QP data construction, working-set translation, KKT residual formulas):
 *     restated from src/QPhandler.cpp and src/qpOASESInterface.cpp; those files cannot be
 *     compiled here (they need qpOASES + Ipopt), so they are pinned by reading only.
 *   - Active-set QP/LP arithmetic (row D): qpOASES 3.2.1 is an un-vendored third-party
 *     dependency (CMakeLists.txt:81-94) that is absent from /root/reference and from this
 *     container.  "PARITY UNPINNED": the oracle restates the published online active-set
 *     strategy (Ferreau, Bock, Diehl 2008; Ferreau et al. 2014) and is validated by
 *     independent ground truth (exhaustive active-set enumeration on small strictly convex
 *     QPs, and the reference's own KKT acceptance test at 1e-6), not by qpOASES outputs.
 */
#ifndef RESTARTSQP_ORACLE_H
#define RESTARTSQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: include/sqphot/Utils.hpp:35-37 ---- */
#define ORC_INF 1.0e18
#define ORC_M_EPS 1.0e-16
#define ORC_SQRT_M_EPS 1.0e-8

/* ActiveType: include/sqphot/Types.hpp:84-89 */
#define ORC_ACTIVE_ABOVE 1
#define ORC_ACTIVE_BELOW (-1)
#define ORC_ACTIVE_BOTH_SIDE (-99)
#define ORC_INACTIVE 0

/* Exitflag QP codes: include/sqphot/Types.hpp:51-73 */
#define ORC_QP_OPTIMAL 20
#define ORC_QPERROR_INTERNAL_ERROR 21
#define ORC_QPERROR_INFEASIBLE 22
#define ORC_QPERROR_UNBOUNDED 23
#define ORC_QPERROR_EXCEED_MAX_ITER 24
#define ORC_QPERROR_NOTINITIALISED 25
#define ORC_QPERROR_PERFORMINGHOMOTOPY 28
#define ORC_QPERROR_UNKNOWN 30

/* ------------------------------------------------------------------ L0 ---- */

/* Entry-list construction of SpHbMat::setStructure(rhs, I_info), src/SpHbMat.cpp:203-227:
 * the zJ triplets followed by the identity blocks.  Returns total entry count. */
int orc_expand_A(int zJ, const int* row1, const int* col1, const double* val, int I_len,
                 const int* I_irow, const int* I_jcol, const int* I_size, const double* I_value,
                 int* e_row1, int* e_col1, double* e_val);

/* Entry-list construction of SpHbMat::setStructure(rhs), src/SpHbMat.cpp:296-309:
 * every off-diagonal of a symmetric triplet is followed by its mirror. Returns count. */
int orc_expand_H(int zH, const int* row1, const int* col1, const double* val, int symmetric,
                 int* e_row1, int* e_col1, double* e_val);

/* Sort by (col,row) [tie-break: entry counter] and emit 0-based CSC + `order`,
 * src/SpHbMat.cpp:253-264 with comparator include/sqphot/SpHbMat.hpp:370-380. */
void orc_csc_from_entries(int ncol, int z, const int* e_row1, const int* e_col1,
                          const double* e_val, int* colptr, int* rowidx, double* val, int* order);

/* SpHbMat::setMatVal(rhs, I_info): src/SpHbMat.cpp:368-380 */
void orc_setmatval_A(int zJ, const int* order, const double* jac_val, double* csc_val);
/* SpHbMat::setMatVal(rhs): src/SpHbMat.cpp:383-393 */
void orc_setmatval_H(int zH, const int* row1, const int* col1, int symmetric, const int* order,
                     const double* h_val, double* csc_val);

/* SpHbMat::times / transposed_times, CSC branch: src/SpHbMat.cpp:720-736, 679-695 */
void orc_csc_times(int nrow, int ncol, const int* colptr, const int* rowidx, const double* val,
                   const double* x, double* y);
void orc_csc_transposed_times(int nrow, int ncol, const int* colptr, const int* rowidx,
                              const double* val, const double* x, double* y);
/* SpTripletMat::times / transposed_times: src/SpTripletMat.cpp:237-258, 311-323 */
void orc_triplet_times(int nrow, int ncol, int z, const int* row1, const int* col1,
                       const double* val, int symmetric, int transpose, const double* x, double* y);
/* Utils oneNorm/infNorm: src/Utils.cpp:65-83 */
double orc_one_norm(const double* x, int n);
double orc_inf_norm(const double* x, int n);

/* ------------------------------------------------- L2: QPhandler data ---- */
/* QPhandler::set_bounds / update_bounds / update_delta (non-QORE branch):
 * src/QPhandler.cpp:185-201, 358-367, 559-564.  mode: 0=set_bounds, 1=update_bounds
 * (refreshes lbA but NOT ubA: quirk 2), 2=update_delta. */
void orc_qp_bounds(int mode, int n, int m, double delta, const double* x_l, const double* x_u,
                   const double* x_k, const double* c_l, const double* c_u, const double* c_k,
                   double* lb, double* ub, double* lbA, double* ubA);
/* QPhandler::set_g / update_grad / update_penalty: src/QPhandler.cpp:287-292, 461-462, 439-440.
 * grad may be NULL (penalty only); rho<0 means "leave slack entries". */
void orc_qp_g(int n, int m, const double* grad, double rho, double* g);
/* QPhandler::get_infea_measure_model: src/QPhandler.cpp:592-594 */
double orc_infea_measure_model(int n, int m, const double* x_qp);

/* ------------------------------------------------- L1: backend glue ------ */
/* qpOASESInterface::get_working_set: src/qpOASESInterface.cpp:835-895 (raw +1/-1/0 ->
 * ActiveType, including the misplaced-parenthesis behaviour at :874 and :880). */
void orc_translate_working_set(int nV, int nC, const int* raw_b, const int* raw_c, const double* x,
                               const double* Ax, const double* lb, const double* ub,
                               const double* lbA, const double* ubA, int* W_b, int* W_c);
/* qpOASESInterface::test_optimality: src/qpOASESInterface.cpp:498-684.
 * out[0..4] = primal, dual, stationarity, complementarity, KKT_error.  Returns 1 iff
 * KKT_error <= 1e-6 (:673).  H may be absent (colptr NULL): then Hx = 0. */
int orc_kkt_residuals(int nV, int nC, const int* A_colptr, const int* A_rowidx, const double* A_val,
                      const int* H_colptr, const int* H_rowidx, const double* H_val, const double* g,
                      const double* lb, const double* ub, const double* lbA, const double* ubA,
                      const double* x, const double* y, const int* W_b, const int* W_c, double* out);

/* ------------------------------------------------- row D: active-set QP -- */
typedef struct orc_qp orc_qp;

typedef struct {
    int max_iter;          /* nWSR (Options::qp_maxiter / lp_maxiter) */
    int refactor_every;    /* enableCholeskyRefactorisation (setToReliable: 1) */
    int refine_steps;      /* numRefinementSteps */
    int enable_flipping;   /* enableFlippingBounds */
    int enable_drift;      /* enableDriftCorrection */
    int enable_ramping;    /* enableRamping */
} orc_qp_options;

void orc_qp_default_options(orc_qp_options* o);
/* qpOASESInterface::handle_error (src/qpOASESInterface.cpp:686-758): to be called when the last solve did not end OPTIMAL.
 * Returns the new status; *iters_added = iterations the recovery adds to Stats::qp_iter.  force_guess != 0 (test hook) takes
 * the infeasible branch (init from x0 = [0; max(0, lbA); -min(0, ubA)]) whatever the status. */
int orc_qp_handle_error(orc_qp* q, const orc_qp_options* opt, int force_guess, int* iters_added);
/* deviation study (process-global, tests only): flags 1 = true division, 2 = ratio test as num < t*den, 4 = maxDualJump cap;
 * refine_steps = numRefinementSteps.  (0, 0) = the shipped behaviour. */
void orc_qp_set_variant(int flags, int refine_steps);
orc_qp* orc_qp_create(int nV, int nC);
void orc_qp_destroy(orc_qp* q);
/* Cold start ("init"): H may be NULL colptr for an LP.  Returns Exitflag (20 = optimal). */
int orc_qp_init(orc_qp* q, const orc_qp_options* opt, const int* H_colptr, const int* H_rowidx,
                const double* H_val, const double* g, const int* A_colptr, const int* A_rowidx,
                const double* A_val, const double* lb, const double* ub, const double* lbA,
                const double* ubA, int is_lp);
/* Hot start with unchanged matrices ("hotstart(g,lb,ub,lbA,ubA)"). */
int orc_qp_hotstart(orc_qp* q, const orc_qp_options* opt, const double* g, const double* lb,
                    const double* ub, const double* lbA, const double* ubA);
/* Hot start with new matrix values, same pattern ("hotstart(H,g,A,...)"). */
int orc_qp_hotstart_matrices(orc_qp* q, const orc_qp_options* opt, const double* H_val,
                             const double* A_val, const double* g, const double* lb,
                             const double* ub, const double* lbA, const double* ubA);
void orc_qp_get_solution(const orc_qp* q, double* x, double* y, double* obj, int* iters);
void orc_qp_get_working_set(const orc_qp* q, int* raw_b, int* raw_c);
/* oracle_batch.c: OpenMP driver over independent instances (one solver object per thread) */
int orc_max_threads(void);
int orc_qp_solve_batch(int B, int nV, int nC, const int* Hp, const int* Hi, const double* Hv, int Hv_stride,
                       const int* Ap, const int* Ai, const double* Av, int Av_stride, const double* g,
                       const double* lb, const double* ub, const double* lbA, const double* ubA, int is_lp,
                       int max_iter, double* x, double* y, double* obj, int* status, int* iters, int nthreads);
/* flop counter for the roofline model of SURVEY.md section 8(d) */
double orc_qp_get_flops(const orc_qp* q);
int orc_qp_get_max_free(const orc_qp* q);
/* the FIXED <-> VARIED flip of the matrix status (src/qpOASESInterface.cpp:202-207): init(H, g, A, ..., x_qp, y_qp, &bounds), a
 * fresh init whose auxiliary QP is built from the previous solution, the previous bound statuses and the sign of the previous
 * constraint multipliers */
int orc_qp_reinit(orc_qp* q, const orc_qp_options* opt, const double* H_val, const double* A_val, const double* g,
                  const double* lb, const double* ub, const double* lbA, const double* ubA);
/* 1 if the last orc_qp_hotstart_matrices already performed the cold re-init of handle_error (src/qpOASESInterface.cpp:746-754) */
int orc_qp_get_fell_back(const orc_qp* q);

/* ------------------------------------------------- the caller: Sl1QP outer loop (oracle_sqp.c) -- */
/* NLP callbacks for ONE instance (generated as C from the .nl DAG by restartsqp_b200/nl_reader.py):
 * fc: f and c at x;  all: f, c, grad f, Jacobian triplet values, Lagrangian-Hessian triplet values (lam as handed to eval_h). */
typedef void (*orc_nlp_fc)(const double* x, double* f, double* c);
typedef void (*orc_nlp_all)(const double* x, const double* lam, double* f, double* c, double* grad, double* jac, double* hess);
typedef struct {
    int n, m, zJ, zH;
    const int *J_row1, *J_col1, *H_row1, *H_col1; /* 1-based triplet patterns (Jacobian column-major, Hessian upper triangle) */
    const double *x_l, *x_u, *c_l, *c_u;
    orc_nlp_fc fc;
    orc_nlp_all all;
    /* Options, src/Options.cpp:19-57 */
    int iter_max, penalty_update, penalty_iter_max, qp_maxiter, lp_maxiter;
    double eta_c, eta_s, eta_e, gamma_c, gamma_e, delta, delta_min, delta_max, tol, penalty_update_tol, rho, rho_max,
        increase_parm, eps1, eps1_change_parm, eps2, opt_prim_fea_tol, opt_dual_fea_tol, opt_compl_tol, opt_stat_tol;
    int second_order_correction; /* Options::second_order_correction (src/Options.cpp:26), off by default */
} orc_sqp_problem;
/* Algorithm::initialization + Optimize for one starting point (src/Algorithm.cpp:55-168, 438-472). */
int orc_sqp_solve(const orc_sqp_problem* P, const double* x0, const double* lam0, double* x_out, double* f_out,
                  int* exitflag, int* iters, long long* qp_iters, double* kkt_out);
/* test aid: QP solves that failed inside the penalty loop (src/Algorithm.cpp:932-935, 958-961) since the last call */
long long orc_sqp_penalty_qp_failures(void);
/* One solve per instance x0[B][n] on the host cores; returns the number of threads used. */
int orc_sqp_solve_batch(const orc_sqp_problem* P, int B, const double* x0, const double* lam0, double* x, double* f,
                        int* exitflag, int* iters, long long* qp_iters, double* kkt, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
