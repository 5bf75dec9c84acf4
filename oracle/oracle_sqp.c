/* oracle/oracle_sqp.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Single-instance restatement of the Sl1QP outer loop of the reference, src/Algorithm.cpp, as it runs behind
 * test/simple_test.cpp:72-78 with a qpOASES-order backend:
 *
 *   Optimize                    :55-168      setupQP                     :645-697
 *   check_optimality            :170-411     ratio_test                  :722-801
 *   get_trial_point_info        :414-429     update_radius               :820-849
 *   initialization              :438-472     update_penalty_parameter    :886-1028
 *   cal_infea                   :577-602     get_multipliers             :618-630
 *
 * with QPhandler (src/QPhandler.cpp:167-261, 272-297, 342-467, 533-567, 470-499) and the backend's init / hotstart state
 * machine and one-retry recovery (src/qpOASESInterface.cpp:137-284, 686-758, 817-833) on top of the oracle's active-set
 * solver (oracle_qp.c).  The NLP is a pair of callbacks (f,c and f,c,grad,Jacobian,Hessian at one point) that
 * restartsqp_b200/nl_reader.py generates as C from the .nl expression DAG.  orc_sqp_solve_batch runs one solve per instance on
 * the host cores (pthreads): it is the CPU figure bench.py prints beside the GPU's SQP solves/s, and a third, independent
 * statement of the loop next to the numpy mirror (sqp_driver.py) and the device kernels (csrc/sqp_outer.cu).
 *
 * PINNED against the reference's own code: oracle/_ref/algorithm_nl* links src/Algorithm.cpp, src/SQPTNLP.cpp and src/QPhandler.cpp
 * (patched only by integration/restartsqp_cuda_backend.patch) with the QORE-layout plugin; run on the CPU twin of the C ABI it gives
 * this file's exit flags, outer and QP iteration counts, iterates and objectives bit for bit on 23 models x 6 starts
 * (tests/test_reference_algorithm.py; the differences -- the QORE setters' clipping of infinite bounds, failure labels, the
 * reference's acceptance of a NaN KKT error -- are asserted there).  What stays unpinned is the QP solver under it (oracle_qp.c).
 *
 * Deliberate choices shared with the product's drivers: update_bounds refreshes ubA as well (mode 3; the reference's
 * non-QORE branch leaves it stale), a failed QP ends the instance with the QP status as exit flag
 * (the reference throws), second-order correction off (src/Options.cpp:26).
 */
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

enum { EX_OPTIMAL = 0, EX_EXCEED_MAX_ITER = 2, EX_TRUST_REGION_TOO_SMALL = 4, EX_UNKNOWN = -99,
       EX_QP_UNCHANGED = 7 /* not a value of the reference's Exitflag: setupQP throws QP_UNCHANGED there (src/Algorithm.cpp:651-670) */ };
enum { CT_BOUNDED = 5, CT_EQUAL = -5, CT_BOUNDED_ABOVE = 9, CT_BOUNDED_BELOW = 1, CT_UNBOUNDED = 0 };
enum { MS_UNDEFINED = -1, MS_FIXED = 0, MS_VARIED = 1 };

static void* zal(size_t n, size_t s) { return calloc(n > 0 ? n : 1, s); }

/* src/Utils.cpp:29-45 (including its `upper_bound > INF` test for BOUNDED_BELOW) */
static int classify(double lo, double hi) {
    if (lo > -ORC_INF && hi < ORC_INF) return ((hi - lo) < 1.0e-8) ? CT_EQUAL : CT_BOUNDED;
    if (lo > -ORC_INF && hi > ORC_INF) return CT_BOUNDED_BELOW;
    if (hi < ORC_INF && lo < -ORC_INF) return CT_BOUNDED_ABOVE;
    return CT_UNBOUNDED;
}

/* ---------------------------------------------------------------- backend + handler of one QP type */
typedef struct {
    int n, m, nV, nC, is_lp, maxiter;
    int zA, zH;
    int *Ap, *Ai, *Aorder, *Hp, *Hi, *Horder;
    double *Av, *Hv;
    double *g, *lb, *ub, *lbA, *ubA;
    double *x, *y, obj, kkt[5];
    int status, iters;
    orc_qp* solver;
    int inited, first_solved, upd_A, upd_H, old_ms, new_ms, has_A, has_H;
} backend;

static backend* backend_create(int n, int m, int is_lp, int maxiter) {
    backend* b = (backend*)zal(1, sizeof(backend));
    b->n = n; b->m = m; b->nV = n + 2 * m; b->nC = m; b->is_lp = is_lp; b->maxiter = maxiter;
    b->g = zal(b->nV, 8); b->lb = zal(b->nV, 8); b->ub = zal(b->nV, 8); b->lbA = zal(m, 8); b->ubA = zal(m, 8);
    b->x = zal(b->nV, 8); b->y = zal(b->nV + m, 8);
    b->solver = orc_qp_create(b->nV, b->nC);
    b->status = ORC_QPERROR_NOTINITIALISED;
    b->old_ms = b->new_ms = MS_UNDEFINED;
    return b;
}
static void backend_destroy(backend* b) {
    if (!b) return;
    free(b->Ap); free(b->Ai); free(b->Aorder); free(b->Hp); free(b->Hi); free(b->Horder); free(b->Av); free(b->Hv);
    free(b->g); free(b->lb); free(b->ub); free(b->lbA); free(b->ubA); free(b->x); free(b->y);
    orc_qp_destroy(b->solver);
    free(b);
}
/* set_A: first call builds [J I -I] in CSC (src/QPhandler.cpp:41-51, src/SpHbMat.cpp:196-268), later calls refresh values */
static void backend_set_A(backend* b, int zJ, const int* row1, const int* col1, const double* val) {
    if (!b->has_A) {
        int I_irow[2] = {1, 1}, I_jcol[2] = {b->n + 1, b->n + b->m + 1}, I_size[2] = {b->m, b->m};
        double I_val[2] = {1.0, -1.0};
        int z = zJ + 2 * b->m;
        int *er = zal(z, 4), *ec = zal(z, 4);
        double *ev = zal(z, 8), *zero = zal(zJ, 8);
        orc_expand_A(zJ, row1, col1, zero, 2, I_irow, I_jcol, I_size, I_val, er, ec, ev);
        b->zA = z;
        b->Ap = zal(b->nV + 1, 4); b->Ai = zal(z, 4); b->Aorder = zal(z, 4); b->Av = zal(z, 8);
        orc_csc_from_entries(b->nV, z, er, ec, ev, b->Ap, b->Ai, b->Av, b->Aorder);
        free(er); free(ec); free(ev); free(zero);
        b->has_A = 1;
    }
    orc_setmatval_A(zJ, b->Aorder, val, b->Av);
    if (b->first_solved) b->upd_A = 1;
}
static void backend_set_H(backend* b, int zH, const int* row1, const int* col1, const double* val) {
    if (!b->has_H) {
        int *er = zal(2 * zH, 4), *ec = zal(2 * zH, 4);
        double *ev = zal(2 * zH, 8), *zero = zal(zH, 8);
        int z = orc_expand_H(zH, row1, col1, zero, 1, er, ec, ev);
        b->zH = z;
        b->Hp = zal(b->nV + 1, 4); b->Hi = zal(z, 4); b->Horder = zal(z, 4); b->Hv = zal(z, 8);
        orc_csc_from_entries(b->nV, z, er, ec, ev, b->Hp, b->Hi, b->Hv, b->Horder);
        free(er); free(ec); free(ev); free(zero);
        b->has_H = 1;
    }
    orc_setmatval_H(zH, row1, col1, 1, b->Horder, val, b->Hv);
    if (b->first_solved) b->upd_H = 1;
}
/* optimizeQP / optimizeLP + test_optimality */
static void backend_solve(backend* b) {
    orc_qp_options o;
    orc_qp_default_options(&o);
    o.max_iter = b->maxiter;
    int mode = 0; /* 0 cold, 1 fixed, 2 varied, 3 status flip: init from the previous solution */
    if (b->first_solved) {
        int varied = b->upd_A || b->upd_H;
        if (b->old_ms == MS_UNDEFINED) b->old_ms = varied ? MS_VARIED : MS_FIXED;
        else { if (b->new_ms != MS_UNDEFINED) b->old_ms = b->new_ms; b->new_ms = varied ? MS_VARIED : MS_FIXED; }
        if (b->new_ms == MS_UNDEFINED) mode = (b->old_ms == MS_FIXED) ? 1 : 2;
        else if (b->new_ms == MS_FIXED && b->old_ms == MS_FIXED) mode = 1;
        else if (b->new_ms == MS_VARIED && b->old_ms == MS_VARIED) mode = 2;
        else { mode = 3; b->new_ms = b->old_ms = MS_UNDEFINED; }
    }
    const int* Hp = b->is_lp ? NULL : b->Hp; const int* Hi = b->is_lp ? NULL : b->Hi; const double* Hv = b->is_lp ? NULL : b->Hv;
    int st, its = 0, it1 = 0;
    if (mode == 0 || !b->inited) {
        st = orc_qp_init(b->solver, &o, Hp, Hi, Hv, b->g, b->Ap, b->Ai, b->Av, b->lb, b->ub, b->lbA, b->ubA, b->is_lp);
        orc_qp_get_solution(b->solver, NULL, NULL, NULL, &its);
    } else {
        st = (mode == 1) ? orc_qp_hotstart(b->solver, &o, b->g, b->lb, b->ub, b->lbA, b->ubA)
           : (mode == 2) ? orc_qp_hotstart_matrices(b->solver, &o, Hv, b->Av, b->g, b->lb, b->ub, b->lbA, b->ubA)
                         : orc_qp_reinit(b->solver, &o, Hv, b->Av, b->g, b->lb, b->ub, b->lbA, b->ubA);
        orc_qp_get_solution(b->solver, NULL, NULL, NULL, &its);
    }
    if (st != ORC_QP_OPTIMAL) { /* handle_error (:160-162, :217-219), after an init as well as after a hot start */
        st = orc_qp_handle_error(b->solver, &o, 0, &it1);
        its += it1;
    }
    b->inited = (st == ORC_QP_OPTIMAL);
    orc_qp_get_solution(b->solver, b->x, b->y, &b->obj, NULL);
    b->status = st; b->iters = its;
    int *wb = zal(b->nV, 4), *wc = zal(b->nC, 4), *Wb = zal(b->nV, 4), *Wc = zal(b->nC, 4);
    double* Ax = zal(b->nC, 8);
    orc_qp_get_working_set(b->solver, wb, wc);
    orc_csc_times(b->nC, b->nV, b->Ap, b->Ai, b->Av, b->x, Ax);
    orc_translate_working_set(b->nV, b->nC, wb, wc, b->x, Ax, b->lb, b->ub, b->lbA, b->ubA, Wb, Wc);
    orc_kkt_residuals(b->nV, b->nC, b->Ap, b->Ai, b->Av, Hp, Hi, Hv, b->g, b->lb, b->ub, b->lbA, b->ubA, b->x, b->y, Wb, Wc, b->kkt);
    free(wb); free(wc); free(Wb); free(Wc); free(Ax);
    b->upd_A = b->upd_H = 0;
    b->first_solved = 1;
}

/* ---------------------------------------------------------------- the outer loop */
typedef struct {
    const orc_sqp_problem* P;
    double *x_k, *c_k, f_k, *grad, *jac, *hess, *lam_c, *lam_x, *neg_lam;
    double delta, rho, eps1, infea, infea_model, infea_trial, f_trial, actual_red, pred_red, kkt_err;
    double *p_k, *x_trial, *c_trial, *g_new, *j_new, *h_new, *diff;
    int *bound_type, *cons_type;
    int exitflag, iter, pen_trial;
    long long qp_iter;
    int updA, updH, updBounds, updDelta, updPenalty, updG, first, lp_failed;
    backend *qp, *lp;
} sqp;

static double cal_infea(const orc_sqp_problem* P, const double* c) {
    double s = 0.0;
    for (int i = 0; i < P->m; i++) {
        double below = (c[i] < P->c_l[i]) ? P->c_l[i] - c[i] : 0.0;
        double above = (c[i] >= P->c_l[i] && c[i] > P->c_u[i]) ? c[i] - P->c_u[i] : 0.0;
        s = s + below + above;
    }
    return s;
}
static void get_multipliers(sqp* S) {
    const orc_sqp_problem* P = S->P;
    int nV = P->n + 2 * P->m;
    for (int i = 0; i < P->m; i++) S->lam_c[i] = S->qp->y[nV + i];
    for (int i = 0; i < P->n; i++) S->lam_x[i] = S->qp->y[i];
}
static void check_optimality(sqp* S) {
    const orc_sqp_problem* P = S->P;
    int n = P->n, m = P->m;
    const double *mv = S->lam_x, *mc = S->lam_c;
    double primal = S->infea, dual = 0.0, compl_ = 0.0, stat = 0.0;
    for (int i = 0; i < n; i++) {
        int t = S->bound_type[i];
        dual = dual + (t == CT_BOUNDED_ABOVE ? fmax(mv[i], 0.0) : 0.0) + (t == CT_BOUNDED_BELOW ? -fmin(mv[i], 0.0) : 0.0);
    }
    for (int i = 0; i < m; i++) {
        int t = S->cons_type[i];
        dual = dual + (t == CT_BOUNDED_ABOVE ? fmax(mc[i], 0.0) : 0.0) + (t == CT_BOUNDED_BELOW ? -fmin(mc[i], 0.0) : 0.0);
    }
    for (int i = 0; i < m; i++) {
        int t = S->cons_type[i];
        compl_ = compl_ + (t == CT_BOUNDED_ABOVE ? fabs(mc[i] * (P->c_u[i] - S->c_k[i])) : 0.0)
                        + (t == CT_BOUNDED_BELOW ? fabs(mc[i] * (S->c_k[i] - P->c_l[i])) : 0.0)
                        + (t == CT_UNBOUNDED ? fabs(mc[i]) : 0.0);
    }
    for (int i = 0; i < n; i++) {
        int t = S->bound_type[i];
        compl_ = compl_ + (t == CT_BOUNDED_ABOVE ? fabs(mv[i] * (P->x_u[i] - S->x_k[i])) : 0.0)
                        + (t == CT_BOUNDED_BELOW ? fabs(mv[i] * (S->x_k[i] - P->x_l[i])) : 0.0)
                        + (t == CT_UNBOUNDED ? fabs(mv[i]) : 0.0);
    }
    for (int i = 0; i < n; i++) S->diff[i] = 0.0;
    for (int k = 0; k < P->zJ; k++) S->diff[P->J_col1[k] - 1] += S->jac[k] * mc[P->J_row1[k] - 1];
    for (int i = 0; i < n; i++) { double d = S->diff[i] + mv[i] - S->grad[i]; stat = stat + fabs(d); }
    S->kkt_err = dual + primal + compl_ + stat;
    if (primal < P->opt_prim_fea_tol && dual < P->opt_dual_fea_tol && compl_ < P->opt_compl_tol && stat < P->opt_stat_tol) S->exitflag = EX_OPTIMAL;
}
/* QPhandler::solveQP + the failure handling of Algorithm::Optimize; returns 1 if the QP was solved and accepted */
static int solve_qp(sqp* S) {
    backend_solve(S->qp);
    S->qp_iter += S->qp->iters;
    int ok = (S->qp->kkt[4] <= 1.0e-6) && S->qp->status == ORC_QP_OPTIMAL;
    if (!ok) S->exitflag = (S->qp->status == ORC_QP_OPTIMAL) ? ORC_QPERROR_INTERNAL_ERROR : S->qp->status;
    return ok;
}
static double slack_norm(const orc_sqp_problem* P, const double* x) {
    double s = 0.0;
    for (int i = P->n; i < P->n + 2 * P->m; i++) s = s + fabs(x[i]);
    return s;
}

static void setup_qp(sqp* S) {
    const orc_sqp_problem* P = S->P;
    backend* q = S->qp;
    if (S->first) {
        backend_set_A(q, P->zJ, P->J_row1, P->J_col1, S->jac);
        backend_set_H(q, P->zH, P->H_row1, P->H_col1, S->hess);
        orc_qp_bounds(0, P->n, P->m, S->delta, P->x_l, P->x_u, S->x_k, P->c_l, P->c_u, S->c_k, q->lb, q->ub, q->lbA, q->ubA);
        orc_qp_g(P->n, P->m, S->grad, S->rho, q->g);
        S->first = 0;
        return;
    }
    /* no Update_* flag raised since the last solve: the reference throws QP_UNCHANGED (src/Algorithm.cpp:651-670), which nothing
     * catches; the run ends here with its own exit flag instead of re-solving the same QP until iter_max */
    if (!(S->updA || S->updH || S->updBounds || S->updDelta || S->updPenalty || S->updG)) { S->exitflag = EX_QP_UNCHANGED; return; }
    if (S->updA) backend_set_A(q, P->zJ, P->J_row1, P->J_col1, S->jac);
    if (S->updH) backend_set_H(q, P->zH, P->H_row1, P->H_col1, S->hess);
    if (S->updBounds) orc_qp_bounds(3, P->n, P->m, S->delta, P->x_l, P->x_u, S->x_k, P->c_l, P->c_u, S->c_k, q->lb, q->ub, q->lbA, q->ubA);
    else if (S->updDelta) orc_qp_bounds(2, P->n, P->m, S->delta, P->x_l, P->x_u, S->x_k, NULL, NULL, NULL, q->lb, q->ub, q->lbA, q->ubA);
    if (S->updPenalty) orc_qp_g(P->n, P->m, NULL, S->rho, q->g);
    if (S->updG) orc_qp_g(P->n, P->m, S->grad, -1.0, q->g);
    S->updA = S->updH = S->updBounds = S->updDelta = S->updPenalty = S->updG = 0;
}

/* test aid (process-global): how many QP solves failed inside the penalty loop since the last call */
static long long g_pen_qp_failures = 0;
long long orc_sqp_penalty_qp_failures(void) { long long v = g_pen_qp_failures; g_pen_qp_failures = 0; return v; }

static void update_penalty_parameter(sqp* S) {
    const orc_sqp_problem* P = S->P;
    if (!P->penalty_update) return;
    S->infea_model = slack_norm(P, S->qp->x);
    if (!(S->infea_model > P->penalty_update_tol)) return;
    const double infea_model_tmp = S->infea_model;
    double rho_trial = S->rho;
    backend* lp = S->lp;
    orc_qp_bounds(0, P->n, P->m, S->delta, P->x_l, P->x_u, S->x_k, P->c_l, P->c_u, S->c_k, lp->lb, lp->ub, lp->lbA, lp->ubA);
    orc_qp_g(P->n, P->m, NULL, S->rho, lp->g);
    backend_set_A(lp, P->zJ, P->J_row1, P->J_col1, S->jac);
    backend_solve(lp);
    S->qp_iter += lp->iters;
    if (lp->status != ORC_QP_OPTIMAL) { S->exitflag = lp->status; S->lp_failed = 1; return; } /* LP_NOT_OPTIMAL leaves Optimize, :900-906 */
    const double infea_infty = slack_norm(P, lp->x);
    const int feasible_lp = infea_infty <= P->penalty_update_tol;
    int need = 1;
    for (;;) {
        int cont_a = feasible_lp && S->infea_model > P->penalty_update_tol && rho_trial < P->rho_max;
        int cont_b = !feasible_lp && ((S->infea - S->infea_model) < S->eps1 * (S->infea - infea_infty)) &&
                     S->pen_trial < P->penalty_iter_max && rho_trial < P->rho_max;
        if (!((cont_a || cont_b) && S->exitflag == EX_UNKNOWN)) break;
        rho_trial = fmin(P->rho_max, rho_trial * P->increase_parm);
        S->pen_trial += 1;
        orc_qp_g(P->n, P->m, NULL, rho_trial, S->qp->g);
        if (solve_qp(S)) S->infea_model = slack_norm(P, S->qp->x);
        else __sync_fetch_and_add(&g_pen_qp_failures, 1);
        /* else: QP_NOT_OPTIMAL inside the loop only leaves the loop (:932-935, :958-961); the acceptance test below then sees
         * the objective of an unsolved QP (INFTY) and takes its failure branch, and Optimize runs the rest of the iteration
         * (trial point, ratio test, iter++, check_optimality) before its loop condition ends the solve */
    }
    if (need && rho_trial > S->rho) {
        const double qp_obj = S->qp->obj;
        if (rho_trial * S->infea - qp_obj >= P->eps2 * rho_trial * (S->infea - S->infea_model)) {
            S->eps1 += (1 - S->eps1) * P->eps1_change_parm;
            for (int i = 0; i < P->n; i++) S->p_k[i] = S->qp->x[i];
            S->rho = rho_trial;
        } else {
            S->infea_model = infea_model_tmp;
            S->updPenalty = 1;
        }
    }
}

int orc_sqp_solve(const orc_sqp_problem* P, const double* x0, const double* lam0, double* x_out, double* f_out,
                  int* exitflag, int* iters, long long* qp_iters, double* kkt_out) {
    const int n = P->n, m = P->m;
    sqp S;
    memset(&S, 0, sizeof S);
    S.P = P;
    S.x_k = zal(n, 8); S.c_k = zal(m, 8); S.grad = zal(n, 8); S.jac = zal(P->zJ, 8); S.hess = zal(P->zH, 8);
    S.lam_c = zal(m, 8); S.lam_x = zal(n, 8); S.neg_lam = zal(m, 8);
    S.p_k = zal(n, 8); S.x_trial = zal(n, 8); S.c_trial = zal(m, 8); S.g_new = zal(n, 8); S.j_new = zal(P->zJ, 8); S.h_new = zal(P->zH, 8);
    S.diff = zal(n, 8); S.bound_type = zal(n, 4); S.cons_type = zal(m, 4);
    S.qp = backend_create(n, m, 0, P->qp_maxiter);
    S.lp = backend_create(n, m, 1, P->lp_maxiter);
    /* initialization(), :438-472 */
    S.delta = P->delta; S.rho = P->rho; S.eps1 = P->eps1;
    for (int i = 0; i < n; i++) S.x_k[i] = fmin(fmax(x0[i], P->x_l[i]), P->x_u[i]); /* shift_starting_point, src/SQPTNLP.cpp:140-153 */
    for (int i = 0; i < m; i++) { S.lam_c[i] = lam0 ? lam0[i] : 0.0; S.neg_lam[i] = -S.lam_c[i]; }
    P->all(S.x_k, S.neg_lam, &S.f_k, S.c_k, S.grad, S.jac, S.hess);
    for (int i = 0; i < n; i++) S.bound_type[i] = classify(P->x_l[i], P->x_u[i]);
    for (int i = 0; i < m; i++) S.cons_type[i] = classify(P->c_l[i], P->c_u[i]);
    S.infea = cal_infea(P, S.c_k);
    S.exitflag = EX_UNKNOWN; S.first = 1; S.kkt_err = INFINITY;
    double f_tmp;
    double* c_tmp = zal(m, 8);
    while (S.iter < P->iter_max && S.exitflag == EX_UNKNOWN) {
        setup_qp(&S);
        if (S.exitflag != EX_UNKNOWN) break;
        if (!solve_qp(&S)) break;
        for (int i = 0; i < n; i++) S.p_k[i] = S.qp->x[i];
        update_penalty_parameter(&S);
        if (S.lp_failed) break;
        /* get_trial_point_info */
        double norm_p = 0.0;
        for (int i = 0; i < n; i++) { S.x_trial[i] = S.x_k[i] + S.p_k[i]; norm_p = fmax(norm_p, fabs(S.p_k[i])); }
        P->fc(S.x_trial, &S.f_trial, S.c_trial);
        S.infea_trial = cal_infea(P, S.c_trial);
        /* ratio_test */
        const double P1_x = S.f_k + S.rho * S.infea, P1_t = S.f_trial + S.rho * S.infea_trial;
        S.actual_red = P1_x - P1_t;
        S.pred_red = S.rho * S.infea - S.qp->obj;
        if (S.actual_red >= P->eta_s * S.pred_red && S.actual_red >= -P->tol) {
            S.infea = S.infea_trial; S.f_k = S.f_trial;
            memcpy(S.x_k, S.x_trial, sizeof(double) * n);
            memcpy(S.c_k, S.c_trial, sizeof(double) * m);
            get_multipliers(&S);
            for (int i = 0; i < m; i++) S.neg_lam[i] = -S.lam_c[i];
            P->all(S.x_k, S.neg_lam, &f_tmp, c_tmp, S.grad, S.jac, S.hess);
            S.updA = S.updH = S.updBounds = S.updG = 1;
        } else if (P->second_order_correction && S.exitflag == EX_UNKNOWN) {
            /* second_order_correction (src/Algorithm.cpp:1140-1211): the QP again around the trial point with the gradient
             * H_k p_k + g_k; its solution s_k is added to p_k and the ratio test repeated with the SOC QP's own objective as the
             * predicted reduction (ratio_test reads get_obj_QP(), :728); p_k and the QP data are restored otherwise. */
            backend* q = S.qp;
            memcpy(S.g_new, S.p_k, sizeof(double) * n); /* p_k_tmp (g_new is free until the next accepted step) */
            for (int i = 0; i < n; i++) S.diff[i] = 0.0;
            for (int k = 0; k < P->zH; k++) { /* symmetric-half triplet product in storage order, src/SpTripletMat.cpp:237-258 */
                const int i = P->H_row1[k] - 1, j = P->H_col1[k] - 1;
                S.diff[i] += S.hess[k] * S.p_k[j];
                if (i != j) S.diff[j] += S.hess[k] * S.p_k[i];
            }
            for (int i = 0; i < n; i++) S.diff[i] = S.diff[i] + S.grad[i];
            orc_qp_g(P->n, P->m, S.diff, -1.0, q->g);
            orc_qp_bounds(3, P->n, P->m, S.delta, P->x_l, P->x_u, S.x_trial, P->c_l, P->c_u, S.c_trial, q->lb, q->ub, q->lbA, q->ubA);
            int accepted = 0;
            if (solve_qp(&S)) {
                for (int i = 0; i < n; i++) { S.p_k[i] = S.p_k[i] + q->x[i]; S.x_trial[i] = S.x_k[i] + S.p_k[i]; }
                P->fc(S.x_trial, &S.f_trial, S.c_trial);
                S.infea_trial = cal_infea(P, S.c_trial);
                const double Q1_x = S.f_k + S.rho * S.infea, Q1_t = S.f_trial + S.rho * S.infea_trial;
                S.actual_red = Q1_x - Q1_t;
                S.pred_red = S.rho * S.infea - q->obj;
                accepted = S.actual_red >= P->eta_s * S.pred_red && S.actual_red >= -P->tol;
            }
            if (!accepted) memcpy(S.p_k, S.g_new, sizeof(double) * n);
            /* QP data back to the current point (the reference does this after a rejected correction; after an accepted one the
             * next setupQP rewrites gradient and bounds anyway) */
            orc_qp_g(P->n, P->m, S.grad, -1.0, q->g);
            orc_qp_bounds(3, P->n, P->m, S.delta, P->x_l, P->x_u, S.x_k, P->c_l, P->c_u, S.c_k, q->lb, q->ub, q->lbA, q->ubA);
            if (accepted) {
                /* update_radius reads p_k_->getInfNorm() (src/Algorithm.cpp:822): after an accepted correction that is the norm of the
                 * corrected step p_k + s_k (found by running the reference's own code: tests/test_reference_algorithm.py) */
                norm_p = 0.0;
                for (int i = 0; i < n; i++) norm_p = fmax(norm_p, fabs(S.p_k[i]));
                S.infea = S.infea_trial; S.f_k = S.f_trial;
                memcpy(S.x_k, S.x_trial, sizeof(double) * n);
                memcpy(S.c_k, S.c_trial, sizeof(double) * m);
                get_multipliers(&S);
                for (int i = 0; i < m; i++) S.neg_lam[i] = -S.lam_c[i];
                P->all(S.x_k, S.neg_lam, &f_tmp, c_tmp, S.grad, S.jac, S.hess);
                S.updA = S.updH = S.updBounds = S.updG = 1;
            }
        }
        S.iter += 1;
        get_multipliers(&S);
        check_optimality(&S);
        if (S.exitflag != EX_UNKNOWN) break;
        /* update_radius */
        const int shrink = S.actual_red < P->eta_c * S.pred_red;
        const int grow = !shrink && S.actual_red > P->eta_e * S.pred_red && P->tol > fabs(S.delta - norm_p);
        if (shrink) S.delta = P->gamma_c * S.delta;
        if (grow) S.delta = fmin(P->gamma_e * S.delta, P->delta_max);
        if (shrink || grow) S.updDelta = 1;
        if (S.delta < P->delta_min) { S.exitflag = EX_TRUST_REGION_TOO_SMALL; check_optimality(&S); }
    }
    if (S.iter == P->iter_max && S.exitflag == EX_UNKNOWN) S.exitflag = EX_EXCEED_MAX_ITER;
    if (x_out) memcpy(x_out, S.x_k, sizeof(double) * n);
    if (f_out) *f_out = S.f_k;
    if (exitflag) *exitflag = S.exitflag;
    if (iters) *iters = S.iter;
    if (qp_iters) *qp_iters = S.qp_iter;
    if (kkt_out) *kkt_out = S.kkt_err;
    free(S.x_k); free(S.c_k); free(S.grad); free(S.jac); free(S.hess); free(S.lam_c); free(S.lam_x); free(S.neg_lam);
    free(S.p_k); free(S.x_trial); free(S.c_trial); free(S.g_new); free(S.j_new); free(S.h_new); free(S.diff);
    free(S.bound_type); free(S.cons_type); free(c_tmp);
    backend_destroy(S.qp); backend_destroy(S.lp);
    return 0;
}

/* ---------------------------------------------------------------- batch over the host cores */
typedef struct {
    const orc_sqp_problem* P;
    int B, next;
    pthread_mutex_t mu;
    const double *x0, *lam0;
    double *x, *f, *kkt;
    int *exitflag, *iters;
    long long* qp_iters;
} sqp_batch;

static void* sqp_worker(void* arg) {
    sqp_batch* J = (sqp_batch*)arg;
    const int n = J->P->n;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int b0 = J->next;
        J->next += 16;
        pthread_mutex_unlock(&J->mu);
        if (b0 >= J->B) break;
        for (int b = b0; b < b0 + 16 && b < J->B; b++)
            orc_sqp_solve(J->P, J->x0 + (size_t)b * n, J->lam0, J->x + (size_t)b * n, J->f + b, J->exitflag + b, J->iters + b,
                          J->qp_iters + b, J->kkt ? J->kkt + b : NULL);
    }
    return NULL;
}

int orc_sqp_solve_batch(const orc_sqp_problem* P, int B, const double* x0, const double* lam0, double* x, double* f,
                        int* exitflag, int* iters, long long* qp_iters, double* kkt, int nthreads) {
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > B) nthreads = B;
    if (nthreads < 1) nthreads = 1;
    sqp_batch J;
    J.P = P; J.B = B; J.next = 0; J.x0 = x0; J.lam0 = lam0; J.x = x; J.f = f; J.kkt = kkt; J.exitflag = exitflag; J.iters = iters; J.qp_iters = qp_iters;
    pthread_mutex_init(&J.mu, NULL);
    pthread_t* th = (pthread_t*)zal(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, sqp_worker, &J);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&J.mu);
    return nthreads;
}
