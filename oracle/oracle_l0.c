/* oracle/oracle_l0.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Restatement of the reference's L0 linear algebra and of the L1/L2 glue formulas.
 * Every function cites the reference file:line it follows.  Sums are accumulated in the
 * same order as the reference loops so that results are bit-identical with the reference
 * build in oracle/_ref (both compiled without FMA contraction).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/SpHbMat.cpp:203-227 */
int orc_expand_A(int zJ, const int* row1, const int* col1, const double* val, int I_len,
                 const int* I_irow, const int* I_jcol, const int* I_size, const double* I_value,
                 int* e_row1, int* e_col1, double* e_val) {
    int c = 0;
    for (int i = 0; i < zJ; i++) {
        e_row1[c] = row1[i];
        e_col1[c] = col1[i];
        if (e_val) e_val[c] = val ? val[i] : 0.0;
        c++;
    }
    for (int b = 0; b < I_len; b++)
        for (int j = 0; j < I_size[b]; j++) {
            e_row1[c] = I_irow[b] + j;
            e_col1[c] = I_jcol[b] + j;
            if (e_val) e_val[c] = I_value[b];
            c++;
        }
    return c;
}

/* src/SpHbMat.cpp:296-309 */
int orc_expand_H(int zH, const int* row1, const int* col1, const double* val, int symmetric,
                 int* e_row1, int* e_col1, double* e_val) {
    int c = 0;
    for (int i = 0; i < zH; i++) {
        e_row1[c] = row1[i];
        e_col1[c] = col1[i];
        if (e_val) e_val[c] = val ? val[i] : 0.0;
        c++;
        if (symmetric && row1[i] != col1[i]) {
            e_row1[c] = col1[i];
            e_col1[c] = row1[i];
            if (e_val) e_val[c] = val ? val[i] : 0.0;
            c++;
        }
    }
    return c;
}

typedef struct { int row, col, cnt; } orc_key;

/* include/sqphot/SpHbMat.hpp:370-380 (column-compressed rule), made total with the
 * entry counter as last key (stable tie-break). */
static int key_cmp(const void* a, const void* b) {
    const orc_key* l = (const orc_key*)a;
    const orc_key* r = (const orc_key*)b;
    if (l->col != r->col) return l->col < r->col ? -1 : 1;
    if (l->row != r->row) return l->row < r->row ? -1 : 1;
    return l->cnt < r->cnt ? -1 : (l->cnt > r->cnt ? 1 : 0);
}

/* src/SpHbMat.cpp:253-264 */
void orc_csc_from_entries(int ncol, int z, const int* e_row1, const int* e_col1,
                          const double* e_val, int* colptr, int* rowidx, double* val, int* order) {
    orc_key* k = (orc_key*)malloc(sizeof(orc_key) * (z > 0 ? z : 1));
    for (int i = 0; i < z; i++) { k[i].row = e_row1[i]; k[i].col = e_col1[i]; k[i].cnt = i; }
    qsort(k, z, sizeof(orc_key), key_cmp);
    for (int j = 0; j <= ncol; j++) colptr[j] = 0;
    for (int i = 0; i < z; i++) {
        if (val) val[i] = e_val ? e_val[k[i].cnt] : 0.0;
        rowidx[i] = k[i].row - 1;
        order[k[i].cnt] = i;
        /* reference: for (j = col(1-based); j < ColNum_; j++) ColIndex_[j]++;  (:260-262) */
        for (int j = k[i].col; j < ncol; j++) colptr[j]++;
    }
    colptr[ncol] = z; /* :264 */
    free(k);
}

/* src/SpHbMat.cpp:368-380 */
void orc_setmatval_A(int zJ, const int* order, const double* jac_val, double* csc_val) {
    for (int i = 0; i < zJ; i++) csc_val[order[i]] = jac_val[i];
}

/* src/SpHbMat.cpp:383-393 */
void orc_setmatval_H(int zH, const int* row1, const int* col1, int symmetric, const int* order,
                     const double* h_val, double* csc_val) {
    int j = 0;
    for (int i = 0; i < zH; i++) {
        csc_val[order[j]] = h_val[i];
        j++;
        if (symmetric && col1[i] != row1[i]) {
            csc_val[order[j]] = h_val[i];
            j++;
        }
    }
}

/* src/SpHbMat.cpp:720-736: walk entries in storage order, advancing the column. */
void orc_csc_times(int nrow, int ncol, const int* colptr, const int* rowidx, const double* val,
                   const double* x, double* y) {
    for (int i = 0; i < nrow; i++) y[i] = 0.0;
    for (int c = 0; c < ncol; c++)
        for (int e = colptr[c]; e < colptr[c + 1]; e++) y[rowidx[e]] += val[e] * x[c];
}

/* src/SpHbMat.cpp:679-695, include/sqphot/SpHbMat.hpp:140-177 */
void orc_csc_transposed_times(int nrow, int ncol, const int* colptr, const int* rowidx,
                              const double* val, const double* x, double* y) {
    (void)nrow;
    for (int c = 0; c < ncol; c++) {
        double s = 0.0;
        for (int e = colptr[c]; e < colptr[c + 1]; e++) s += val[e] * x[rowidx[e]];
        y[c] = s;
    }
}

/* src/SpTripletMat.cpp:237-258, 311-323 */
void orc_triplet_times(int nrow, int ncol, int z, const int* row1, const int* col1,
                       const double* val, int symmetric, int transpose, const double* x, double* y) {
    if (symmetric || !transpose) {
        for (int i = 0; i < nrow; i++) y[i] = 0.0;
        for (int i = 0; i < z; i++) {
            y[row1[i] - 1] += val[i] * x[col1[i] - 1];
            if (symmetric && row1[i] != col1[i]) y[col1[i] - 1] += val[i] * x[row1[i] - 1];
        }
    } else {
        for (int i = 0; i < ncol; i++) y[i] = 0.0;
        for (int i = 0; i < z; i++) y[col1[i] - 1] += val[i] * x[row1[i] - 1];
    }
}

/* src/Utils.cpp:65-72 */
double orc_one_norm(const double* x, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) {
        if (x[i] < 0) s -= x[i]; else s += x[i];
    }
    return s;
}

/* src/Utils.cpp:74-83 */
double orc_inf_norm(const double* x, int n) {
    double m = 0;
    for (int i = 0; i < n; i++) {
        double a = x[i] < 0 ? -x[i] : x[i];
        if (a > m) m = a;
    }
    return m;
}

/* src/QPhandler.cpp:185-201 (set_bounds), :358-367 (update_bounds), :559-564 (update_delta) */
void orc_qp_bounds(int mode, int n, int m, double delta, const double* x_l, const double* x_u,
                   const double* x_k, const double* c_l, const double* c_u, const double* c_k,
                   double* lb, double* ub, double* lbA, double* ubA) {
    if (mode == 0 || mode == 1 || mode == 3) {
        for (int i = 0; i < m; i++) {
            lbA[i] = c_l[i] - c_k[i];
            if (mode != 1) ubA[i] = c_u[i] - c_k[i]; /* mode 1: update_bounds never refreshes ubA; mode 3: QORE-branch behaviour */
        }
    }
    for (int i = 0; i < n; i++) {
        double a = x_l[i] - x_k[i], b = x_u[i] - x_k[i];
        lb[i] = a > -delta ? a : -delta; /* std::max(x_l - x_k, -delta) */
        ub[i] = b < delta ? b : delta;   /* std::min(x_u - x_k,  delta) */
    }
    if (mode == 0)
        for (int i = 0; i < 2 * m; i++) ub[n + i] = ORC_INF; /* slack lb stays at its zero-init */
}

/* src/QPhandler.cpp:287-292, 439-440, 461-462 */
void orc_qp_g(int n, int m, const double* grad, double rho, double* g) {
    if (grad) for (int i = 0; i < n; i++) g[i] = grad[i];
    if (rho >= 0) for (int i = n; i < n + 2 * m; i++) g[i] = rho;
}

/* src/QPhandler.cpp:592-594 */
double orc_infea_measure_model(int n, int m, const double* x_qp) {
    return orc_one_norm(x_qp + n, 2 * m);
}

/* src/qpOASESInterface.cpp:846-892 */
void orc_translate_working_set(int nV, int nC, const int* raw_b, const int* raw_c, const double* x,
                               const double* Ax, const double* lb, const double* ub,
                               const double* lbA, const double* ubA, int* W_b, int* W_c) {
    for (int i = 0; i < nV; i++) {
        switch (raw_b[i]) {
        case 1:
            W_b[i] = fabs(x[i] - lb[i]) < ORC_SQRT_M_EPS ? ORC_ACTIVE_BOTH_SIDE : ORC_ACTIVE_ABOVE;
            break;
        case -1:
            W_b[i] = fabs(x[i] - ub[i]) < ORC_SQRT_M_EPS ? ORC_ACTIVE_BOTH_SIDE : ORC_ACTIVE_BELOW;
            break;
        default:
            W_b[i] = ORC_INACTIVE;
        }
    }
    for (int i = 0; i < nC; i++) {
        switch (raw_c[i]) {
        case 1: /* :874  fabs(Ax-lbA<sqrt_m_eps): the comparison is inside fabs */
            W_c[i] = fabs((double)(Ax[i] - lbA[i] < ORC_SQRT_M_EPS)) != 0.0 ? ORC_ACTIVE_BOTH_SIDE
                                                                            : ORC_ACTIVE_ABOVE;
            break;
        case -1: /* :880 */
            W_c[i] = fabs((double)(Ax[i] - ubA[i] < ORC_SQRT_M_EPS)) != 0.0 ? ORC_ACTIVE_BOTH_SIDE
                                                                            : ORC_ACTIVE_BELOW;
            break;
        default:
            W_c[i] = ORC_INACTIVE;
        }
    }
}

static double dmax(double a, double b) { return a > b ? a : b; } /* std::max: (a<b)?b:a */
static double dmin(double a, double b) { return b < a ? b : a; } /* std::min */

/* src/qpOASESInterface.cpp:498-684 */
int orc_kkt_residuals(int nV, int nC, const int* A_colptr, const int* A_rowidx, const double* A_val,
                      const int* H_colptr, const int* H_rowidx, const double* H_val, const double* g,
                      const double* lb, const double* ub, const double* lbA, const double* ubA,
                      const double* x, const double* y, const int* W_b, const int* W_c, double* out) {
    double primal = 0.0, dual = 0.0, compl = 0.0, stat = 0.0;
    double* Ax = (double*)calloc(nC > 0 ? nC : 1, sizeof(double));
    double* gap = (double*)calloc(nV > 0 ? nV : 1, sizeof(double));
    double* Hx = (double*)calloc(nV > 0 ? nV : 1, sizeof(double));
    /* primal feasibility :518-528 */
    for (int i = 0; i < nV; i++) {
        primal += dmax(0.0, lb[i] - x[i]);
        primal += -dmin(0.0, ub[i] - x[i]);
    }
    orc_csc_times(nC, nV, A_colptr, A_rowidx, A_val, x, Ax);
    for (int i = 0; i < nC; i++) {
        primal += dmax(0.0, lbA[i] - Ax[i]);
        primal += -dmin(0.0, ubA[i] - Ax[i]);
    }
    /* dual feasibility :533-578 */
    for (int i = 0; i < nV; i++) {
        switch (W_b[i]) {
        case ORC_INACTIVE: dual += fabs(y[i]); break;
        case ORC_ACTIVE_BELOW: dual += -dmin(0.0, y[i]); break;
        case ORC_ACTIVE_ABOVE: dual += dmax(0.0, y[i]); break;
        default: break;
        }
    }
    for (int i = 0; i < nC; i++) {
        switch (W_c[i]) {
        case ORC_INACTIVE: dual += fabs(y[i + nV]); break;
        case ORC_ACTIVE_BELOW: dual += -dmin(0.0, y[i + nV]); break;
        case ORC_ACTIVE_ABOVE: dual += dmax(0.0, y[i + nV]); break;
        default: break;
        }
    }
    /* stationarity :595-604: gap = A'y_c ; += y_b ; -= g ; -= Hx ; 1-norm */
    orc_csc_transposed_times(nC, nV, A_colptr, A_rowidx, A_val, y + nV, gap);
    if (H_colptr) orc_csc_times(nV, nV, H_colptr, H_rowidx, H_val, x, Hx);
    for (int i = 0; i < nV; i++) gap[i] += y[i];
    for (int i = 0; i < nV; i++) gap[i] -= g[i];
    for (int i = 0; i < nV; i++) gap[i] -= Hx[i];
    for (int i = 0; i < nV; i++) stat += fabs(gap[i]); /* Vector::getOneNorm */
    /* complementarity :611-658 */
    for (int i = 0; i < nV; i++) {
        switch (W_b[i]) {
        case ORC_INACTIVE: compl += fabs(y[i]); break;
        case ORC_ACTIVE_BELOW: compl += fabs(y[i] * (x[i] - lb[i])); break;
        case ORC_ACTIVE_ABOVE: compl += fabs(y[i] * (ub[i] - x[i])); break;
        default: break;
        }
    }
    for (int i = 0; i < nC; i++) {
        switch (W_c[i]) {
        case ORC_INACTIVE: compl += fabs(y[i + nV]); break;
        case ORC_ACTIVE_BELOW: compl += fabs(y[i + nV] * (Ax[i] - lbA[i])); break;
        case ORC_ACTIVE_ABOVE: compl += fabs(y[i + nV] * (ubA[i] - Ax[i])); break;
        default: break;
        }
    }
    out[0] = primal; out[1] = dual; out[2] = stat; out[3] = compl;
    out[4] = compl + stat + dual + primal; /* :664-665 */
    free(Ax); free(gap); free(Hx);
    return out[4] > 1.0e-6 ? 0 : 1; /* :673 */
}
