/* oracle/oracle_qp.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Row D of SURVEY.md section 8a: the active-set arithmetic that the reference obtains from
 * qpOASES 3.2.1 (third-party, pinned at CMakeLists.txt:81, fetched by
 * cmake/ExternalQPOASES.cmake:2-10, NOT present under /root/reference).  PARITY UNPINNED:
 * this file restates the *published* online active-set strategy (Ferreau, Bock, Diehl,
 * "An online active set strategy to overcome the limitations of explicit MPC", 2008;
 * Ferreau et al., "qpOASES: a parametric active-set algorithm for quadratic programming",
 * Math. Prog. Comp. 2014) as the reference drives it:
 *
 *   call sites   src/qpOASESInterface.cpp:155 (init), :180/:191 (hotstart vectors),
 *                :184/:197 (hotstart with matrices), :231-268 (LP variants), :221-222 (results),
 *                :843-844 (working set), :765 (Options::setToReliable)
 *
 *   - null-space method: TQ factorisation of the active constraints on the free variables,
 *     Cholesky factor of the projected Hessian, recomputed at every working-set change
 *     (setToReliable => enableCholeskyRefactorisation = 1);
 *   - cold start from the all-lower-bounds working set with x = 0, y = 0 (initialStatusBounds
 *     = ST_LOWER), auxiliary QP data relaxed by boundRelaxation = 1e4;
 *   - homotopy with ratio tests in the order: duals of active constraints, duals of fixed
 *     bounds, inactive constraints (lower, upper), free bounds (lower, upper); strict '<'
 *     so the first index wins ties; epsNum = -1e3*EPS, epsDen = 1e3*EPS;
 *   - linear-independence test before every addition, exchange step when dependent;
 *   - flipping bounds when a removal exposes non-positive curvature (epsFlipping = 1e3*EPS);
 *   - drift correction after every step; LPs are regularised with epsRegularisation*I.
 *
 * The value INF = 1e18 of the reference (include/sqphot/Utils.hpp:35) is below qpOASES's
 * own infinity (1e20), so every "infinite" bound of the reference is an ordinary finite bound
 * here as well.
 *
 * The CUDA kernel (restartsqp_b200/csrc/qp_kernel.cuh) implements the same steps with the same
 * tie-breaks; tests compare working sets exactly and x/y/objective to 1e-8 relative.
 */
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define QP_EPS 2.221e-16
#define QP_INFTY 1.0e20
#define QP_ZERO 1.0e-25
#define QP_BOUND_RELAX 1.0e4
#define QP_EPS_NUM (-1.0e3 * QP_EPS)
#define QP_EPS_DEN (1.0e3 * QP_EPS)
#define QP_EPS_FLIP (1.0e3 * QP_EPS)
#define QP_EPS_REG (1.0e3 * QP_EPS)
#define QP_EPS_LI (1.0e5 * QP_EPS)
#define QP_MAX_DUAL_JUMP 1.0e8

/* Deviation study (tests/test_oracle_variants.py, DESIGN.md "deviations from setToReliable"): switches that put back the
 * qpOASES behaviours this restatement leaves out, so that a CPU test can show on which inputs they change the result.
 * Process-global, test use only; 0 = the shipped behaviour (what the CUDA kernel implements). */
#define VAR_TRUE_DIVISION 1   /* x / d instead of the Newton-corrected reciprocal of quot() */
#define VAR_RATIO_MULT 2      /* isBlocking as num < t*den (qpOASES form) instead of num/den < t */
#define VAR_MAX_DUAL_JUMP 4   /* exchange step only accepts ratios below maxDualJump = 1e8 */
static int g_variant = 0;
static int g_refine = 0;      /* numRefinementSteps: rounds of iterative refinement of the step direction */
void orc_qp_set_variant(int flags, int refine_steps) { g_variant = flags; g_refine = refine_steps; }

struct orc_qp {
    int nV, nC;
    int has_H, is_lp;
    double reg;
    int *Hp, *Hi; double* Hv;
    int *Ap, *Ai; double* Av;
    double* Ad; /* dense row-major copy of A (nC x nV) for row access */
    double *g, *lb, *ub, *lbA, *ubA;      /* current homotopy data */
    double *gN, *lbN, *ubN, *lbAN, *ubAN; /* target data */
    double *x, *y, *Ax;
    int *sB, *sC;                          /* -1 lower, 0 inactive/free, +1 upper */
    int nFR, nAC;
    int *FR, *AC, *posFR, *posAC;
    double *Q, *T, *R;                     /* ld = nV */
    /* work */
    double *dx, *dy, *dAx, *t1, *t2, *t3, *w, *a, *yv, *zv, *xiC, *xiB, *dg, *dlb, *dub, *dlbA, *dubA;
    int status, iters, initialised, ramp_offset, max_nFR;
    int fell_back; /* last hotstart_matrices could not keep the working set and ran the cold start itself */
    int last_cold; /* the last solve was an init (cold or from a guess), not a hot start */
    double flops;
    int verbose;
};

void orc_qp_default_options(orc_qp_options* o) {
    o->max_iter = 1000;     /* Options::qp_maxiter, src/Options.cpp:47 */
    o->refactor_every = 1;  /* setToReliable */
    o->refine_steps = 0;
    o->enable_flipping = 1;
    o->enable_drift = 1;
    o->enable_ramping = 1;
}

static void* zalloc(size_t n, size_t s) { return calloc(n > 0 ? n : 1, s); }

orc_qp* orc_qp_create(int nV, int nC) {
    orc_qp* q = (orc_qp*)zalloc(1, sizeof(orc_qp));
    q->nV = nV; q->nC = nC;
    int nT = nV + nC;
    q->g = zalloc(nV, 8); q->lb = zalloc(nV, 8); q->ub = zalloc(nV, 8);
    q->lbA = zalloc(nC, 8); q->ubA = zalloc(nC, 8);
    q->gN = zalloc(nV, 8); q->lbN = zalloc(nV, 8); q->ubN = zalloc(nV, 8);
    q->lbAN = zalloc(nC, 8); q->ubAN = zalloc(nC, 8);
    q->x = zalloc(nV, 8); q->y = zalloc(nT, 8); q->Ax = zalloc(nC, 8);
    q->sB = zalloc(nV, 4); q->sC = zalloc(nC, 4);
    q->FR = zalloc(nV, 4); q->AC = zalloc(nC, 4); q->posFR = zalloc(nV, 4); q->posAC = zalloc(nC, 4);
    q->Q = zalloc((size_t)nV * nV, 8); q->T = zalloc((size_t)nV * nV, 8); q->R = zalloc((size_t)nV * nV, 8);
    q->dx = zalloc(nV, 8); q->dy = zalloc(nT, 8); q->dAx = zalloc(nC, 8);
    q->t1 = zalloc(nT, 8); q->t2 = zalloc(nT, 8); q->t3 = zalloc(nT, 8);
    q->w = zalloc(nT, 8); q->a = zalloc(nT, 8); q->yv = zalloc(nT, 8); q->zv = zalloc(nT, 8);
    q->xiC = zalloc(nT, 8); q->xiB = zalloc(nT, 8);
    q->dg = zalloc(nV, 8); q->dlb = zalloc(nV, 8); q->dub = zalloc(nV, 8);
    q->dlbA = zalloc(nC, 8); q->dubA = zalloc(nC, 8);
    q->Ad = zalloc((size_t)nC * nV, 8);
    q->status = ORC_QPERROR_NOTINITIALISED;
    q->verbose = getenv("ORC_QP_VERBOSE") ? atoi(getenv("ORC_QP_VERBOSE")) : 0;
    return q;
}

void orc_qp_destroy(orc_qp* q) {
    if (!q) return;
    free(q->Hp); free(q->Hi); free(q->Hv); free(q->Ap); free(q->Ai); free(q->Av); free(q->Ad);
    free(q->g); free(q->lb); free(q->ub); free(q->lbA); free(q->ubA);
    free(q->gN); free(q->lbN); free(q->ubN); free(q->lbAN); free(q->ubAN);
    free(q->x); free(q->y); free(q->Ax); free(q->sB); free(q->sC);
    free(q->FR); free(q->AC); free(q->posFR); free(q->posAC); free(q->Q); free(q->T); free(q->R);
    free(q->dx); free(q->dy); free(q->dAx); free(q->t1); free(q->t2); free(q->t3); free(q->w);
    free(q->a); free(q->yv); free(q->zv); free(q->xiC); free(q->xiB);
    free(q->dg); free(q->dlb); free(q->dub); free(q->dlbA); free(q->dubA);
    free(q);
}

/* ------------------------------------------------------------ sparse products */
/* out = (H + reg I) v.  H is symmetric (SymSparseMat, src/qpOASESInterface.cpp:412-416), so row c
 * is read from column c of the CSC arrays; for a symmetric H the terms and their order are those
 * of SpHbMat::times (src/SpHbMat.cpp:729-735). */
static void mulH(const orc_qp* q, const double* v, double* out) {
    int nV = q->nV;
    for (int c = 0; c < nV; c++) {
        double s = 0.0;
        if (q->has_H)
            for (int e = q->Hp[c]; e < q->Hp[c + 1]; e++) s += q->Hv[e] * v[q->Hi[e]];
        if (q->reg != 0.0) s += q->reg * v[c];
        out[c] = s;
    }
}
static void mulA(const orc_qp* q, const double* v, double* out) {
    for (int i = 0; i < q->nC; i++) out[i] = 0.0;
    for (int c = 0; c < q->nV; c++)
        for (int e = q->Ap[c]; e < q->Ap[c + 1]; e++) out[q->Ai[e]] += q->Av[e] * v[c];
}
static void mulAT(const orc_qp* q, const double* yc, double* out) {
    for (int c = 0; c < q->nV; c++) {
        double s = 0.0;
        for (int e = q->Ap[c]; e < q->Ap[c + 1]; e++) s += q->Av[e] * yc[q->Ai[e]];
        out[c] = s;
    }
}
static void build_dense_A(orc_qp* q) {
    memset(q->Ad, 0, sizeof(double) * (size_t)q->nC * q->nV);
    for (int c = 0; c < q->nV; c++)
        for (int e = q->Ap[c]; e < q->Ap[c + 1]; e++) q->Ad[(size_t)q->Ai[e] * q->nV + c] += q->Av[e];
}

/* Quotient x / d from the reciprocal ri = 1/d (formed ahead of a substitution loop, in parallel on the GPU) with one
 * Newton correction: q0 = x*ri, r = x - q0*d (exact, fused), q = q0 + r*ri.  Three dependent operations instead of a full
 * division on the critical path; the result is the correctly rounded quotient except in rare double-rounding cases.
 * (A bare x*ri is not accurate enough: on the rho = 1e8 scaled dumps the homotopy then needs 3x the iterations.) */
static inline double quot(double x, double d, double ri) {
    if (g_variant & VAR_TRUE_DIVISION) return x / d;
    double q0 = x * ri;
    double r = fma(-q0, d, x);
    return fma(r, ri, q0);
}

/* ------------------------------------------------------------ Givens helpers */
/* rotation G with [a b] G = [0 r]:  a' = c a - s b,  b' = s a + c b */
static void givens(double a, double b, double* c, double* s, double* r) {
    if (a == 0.0) { *c = 1.0; *s = 0.0; *r = b; return; }
    double h = sqrt(a * a + b * b);
    *c = b / h; *s = a / h; *r = h;
}
static void rot_cols(double* M, int ld, int row_lo, int row_hi, int j, double c, double s) {
    for (int i = row_lo; i < row_hi; i++) {
        double a = M[(size_t)i * ld + j], b = M[(size_t)i * ld + j + 1];
        M[(size_t)i * ld + j] = c * a - s * b;
        M[(size_t)i * ld + j + 1] = s * a + c * b;
    }
}

/* Chain of rotations that compresses w[0..cnt) into its last entry, rotation j acting on columns (j, j+1).
 * Written with the running sum of squares S_j = w_0^2 + ... + w_j^2 (accumulated left to right) so that every
 * rotation can be formed independently once the prefix sums are known: with j0 the first non-zero entry,
 * a_j = 0 (j < j0), a_j0 = w_j0, a_j = sqrt(S_j) (j > j0); rotation j is the identity while a_j == 0 and
 * c_j = w_{j+1}/h, s_j = a_j/h, h = sqrt(S_{j+1}) afterwards.  Returns the compressed value a_{cnt-1}. */
static double rotation_chain(const double* w, int cnt, double* cs, double* sn) {
    double S = w[0] * w[0];
    double a = w[0];
    for (int j = 0; j + 1 < cnt; j++) {
        double Sn = S + w[j + 1] * w[j + 1];
        if (a == 0.0) { cs[j] = 1.0; sn[j] = 0.0; a = w[j + 1]; }
        else { double h = sqrt(Sn); cs[j] = w[j + 1] / h; sn[j] = a / h; a = h; }
        S = Sn;
    }
    return a;
}

/* ------------------------------------------------------------ projected Cholesky */
/* R'R = Z'(H+regI)_FR,FR Z, full recomputation (qpOASES computeProjectedCholesky under
 * enableCholeskyRefactorisation=1).  Returns 0 on success, 1 + failing pivot otherwise. */
static int recompute_R(orc_qp* q) {
    int nV = q->nV, nFR = q->nFR, nZ = q->nFR - q->nAC;
    double* R = q->R;
    if (nZ <= 0) return 0;
    if (q->is_lp) { /* Z'Z = I: R = sqrt(reg) I */
        for (int a = 0; a < nZ; a++)
            for (int b = 0; b < nZ; b++) R[(size_t)a * nV + b] = (a == b) ? sqrt(q->reg) : 0.0;
        return 0;
    }
    for (int b = 0; b < nZ; b++) {
        for (int i = 0; i < nV; i++) q->t1[i] = 0.0;
        for (int p = 0; p < nFR; p++) q->t1[q->FR[p]] = q->Q[(size_t)p * nV + b];
        mulH(q, q->t1, q->t2);
        for (int a = 0; a <= b; a++) {
            double s = 0.0;
            for (int p = 0; p < nFR; p++) s += q->Q[(size_t)p * nV + a] * q->t2[q->FR[p]];
            R[(size_t)a * nV + b] = s;
        }
    }
    q->flops += 2.0 * nZ * nZ * nFR / 2 + (double)nZ * nZ * nZ / 3.0;
    /* in-place upper Cholesky (column by column) */
    for (int j = 0; j < nZ; j++) {
        for (int i = 0; i < j; i++) {
            double s = R[(size_t)i * nV + j];
            for (int k = 0; k < i; k++) s -= R[(size_t)k * nV + i] * R[(size_t)k * nV + j];
            R[(size_t)i * nV + j] = s / R[(size_t)i * nV + i];
        }
        double d = R[(size_t)j * nV + j];
        for (int k = 0; k < j; k++) d -= R[(size_t)k * nV + j] * R[(size_t)k * nV + j];
        if (!(d > QP_ZERO)) return 1 + j;
        R[(size_t)j * nV + j] = sqrt(d);
        for (int i = j + 1; i < nZ; i++) R[(size_t)i * nV + j] = 0.0;
    }
    return 0;
}

/* Border R with the new last null-space column (index nZ-1 after a removal).
 * Returns 1 if the curvature rho2 is acceptable, 0 otherwise (caller flips). */
static int extend_R(orc_qp* q, int check_curvature) {
    int nV = q->nV, nFR = q->nFR, nZ = q->nFR - q->nAC;
    int b = nZ - 1;
    double* R = q->R;
    if (q->is_lp) {
        for (int a = 0; a < b; a++) { R[(size_t)a * nV + b] = 0.0; R[(size_t)b * nV + a] = 0.0; }
        R[(size_t)b * nV + b] = sqrt(q->reg);
        return 1;
    }
    for (int i = 0; i < nV; i++) q->t1[i] = 0.0;
    for (int p = 0; p < nFR; p++) q->t1[q->FR[p]] = q->Q[(size_t)p * nV + b];
    mulH(q, q->t1, q->t2);
    for (int a = 0; a <= b; a++) {
        double s = 0.0;
        for (int p = 0; p < nFR; p++) s += q->Q[(size_t)p * nV + a] * q->t2[q->FR[p]];
        q->w[a] = s;
    }
    /* r = R'^{-1} w[0..b) */
    for (int i = 0; i < b; i++) {
        double s = q->w[i];
        for (int k = 0; k < i; k++) s -= R[(size_t)k * nV + i] * R[(size_t)k * nV + b];
        { double piv = R[(size_t)i * nV + i]; R[(size_t)i * nV + b] = quot(s, piv, 1.0 / piv); }
    }
    double rho2 = q->w[b];
    for (int k = 0; k < b; k++) rho2 -= R[(size_t)k * nV + b] * R[(size_t)k * nV + b];
    q->flops += 2.0 * nZ * nFR + (double)nZ * nZ;
    if (check_curvature ? !(rho2 > QP_EPS_FLIP) : !(rho2 > QP_ZERO)) return 0;
    R[(size_t)b * nV + b] = sqrt(rho2);
    for (int a = 0; a < b; a++) R[(size_t)b * nV + a] = 0.0;
    return 1;
}

/* ------------------------------------------------------------ working-set updates */
/* w = Q' a_FR for constraint row c; returns ||w_Z||^2 and ||a_FR||^2 */
static void constraint_w(orc_qp* q, int c, double* wz2, double* a2) {
    int nV = q->nV, nFR = q->nFR, nZ = q->nFR - q->nAC;
    double s2 = 0.0;
    for (int p = 0; p < nFR; p++) { q->a[p] = q->Ad[(size_t)c * nV + q->FR[p]]; s2 += q->a[p] * q->a[p]; }
    for (int j = 0; j < nFR; j++) {
        double s = 0.0;
        for (int p = 0; p < nFR; p++) s += q->Q[(size_t)p * nV + j] * q->a[p];
        q->w[j] = s;
    }
    double z2 = 0.0;
    for (int j = 0; j < nZ; j++) z2 += q->w[j] * q->w[j];
    *wz2 = z2; *a2 = s2;
    q->flops += 2.0 * nFR * nFR;
}

/* requires w from constraint_w and w_Z != 0 */
static void add_constraint(orc_qp* q, int c, int status) {
    int nV = q->nV, nFR = q->nFR, nAC = q->nAC, nZ = nFR - nAC;
    double r = rotation_chain(q->w, nZ, q->t2, q->t3);
    for (int j = 0; j + 1 < nZ; j++) rot_cols(q->Q, nV, 0, nFR, j, q->t2[j], q->t3[j]);
    for (int j = 0; j < nFR; j++) q->T[(size_t)nAC * nV + j] = (j > nZ - 1) ? q->w[j] : ((j == nZ - 1) ? r : 0.0);
    q->AC[nAC] = c; q->posAC[c] = nAC; q->nAC = nAC + 1; q->sC[c] = status;
    q->flops += 6.0 * nZ * nFR;
}

static void remove_constraint(orc_qp* q, int c) {
    int nV = q->nV, nFR = q->nFR, nAC = q->nAC, k = q->posAC[c];
    double cs, sn, r;
    for (int i = k + 1; i < nAC; i++) {
        int cL = nFR - 1 - i;
        givens(q->T[(size_t)i * nV + cL], q->T[(size_t)i * nV + cL + 1], &cs, &sn, &r);
        rot_cols(q->T, nV, i, nAC, cL, cs, sn);
        q->T[(size_t)i * nV + cL] = 0.0;
        rot_cols(q->Q, nV, 0, nFR, cL, cs, sn);
    }
    for (int i = k + 1; i < nAC; i++) {
        memcpy(q->T + (size_t)(i - 1) * nV, q->T + (size_t)i * nV, sizeof(double) * nFR);
        q->AC[i - 1] = q->AC[i]; q->posAC[q->AC[i - 1]] = i - 1;
    }
    q->nAC = nAC - 1; q->sC[c] = 0; q->posAC[c] = -1;
    q->flops += 6.0 * (nAC - k) * nFR;
}

/* w = row of Q of the free variable v; returns ||w_Z||^2 */
static double bound_w(orc_qp* q, int v) {
    int nV = q->nV, nFR = q->nFR, nZ = q->nFR - q->nAC, p = q->posFR[v];
    double z2 = 0.0;
    for (int j = 0; j < nFR; j++) q->w[j] = q->Q[(size_t)p * nV + j];
    for (int j = 0; j < nZ; j++) z2 += q->w[j] * q->w[j];
    return z2;
}

/* requires w from bound_w and w_Z != 0 */
static void add_bound(orc_qp* q, int v, int status) {
    int nV = q->nV, nFR = q->nFR, nAC = q->nAC, nZ = nFR - nAC, p = q->posFR[v];
    rotation_chain(q->w, nFR, q->t2, q->t3);
    for (int j = 0; j + 1 < nFR; j++) {
        double cs = q->t2[j], sn = q->t3[j];
        rot_cols(q->Q, nV, 0, nFR, j, cs, sn);
        if (j >= nZ - 1) {
            int lo = nFR - 2 - j; if (lo < 0) lo = 0;
            rot_cols(q->T, nV, lo, nAC, j, cs, sn);
        }
    }
    int last = nFR - 1;
    if (p != last) {
        memcpy(q->Q + (size_t)p * nV, q->Q + (size_t)last * nV, sizeof(double) * (nFR - 1));
        q->FR[p] = q->FR[last]; q->posFR[q->FR[p]] = p;
    }
    q->nFR = nFR - 1; q->posFR[v] = -1; q->sB[v] = status;
    q->flops += 6.0 * nFR * nFR;
}

static void remove_bound(orc_qp* q, int v) {
    int nV = q->nV, nFR = q->nFR, nAC = q->nAC;
    double cs, sn, r;
    for (int j = 0; j < nFR; j++) { q->Q[(size_t)nFR * nV + j] = 0.0; q->Q[(size_t)j * nV + nFR] = 0.0; }
    q->Q[(size_t)nFR * nV + nFR] = 1.0;
    for (int i = 0; i < nAC; i++) q->T[(size_t)i * nV + nFR] = q->Ad[(size_t)q->AC[i] * nV + v];
    q->FR[nFR] = v; q->posFR[v] = nFR; nFR++; q->nFR = nFR; q->sB[v] = 0;
    if (nFR > q->max_nFR) q->max_nFR = nFR;
    for (int i = 0; i < nAC; i++) {
        int cL = nFR - 2 - i;
        givens(q->T[(size_t)i * nV + cL], q->T[(size_t)i * nV + cL + 1], &cs, &sn, &r);
        rot_cols(q->T, nV, i, nAC, cL, cs, sn);
        q->T[(size_t)i * nV + cL] = 0.0;
        rot_cols(q->Q, nV, 0, nFR, cL, cs, sn);
    }
    q->flops += 6.0 * nAC * nFR;
}

/* ------------------------------------------------------------ triangular solves with T */
/* T v = b, v indexed by Q column (Y columns nZ..nFR-1); row i has nonzeros at cols >= nFR-1-i */
static void solve_T(const orc_qp* q, const double* b, double* v) {
    int nV = q->nV, nFR = q->nFR, nAC = q->nAC;
    for (int i = 0; i < nAC; i++) {
        int d = nFR - 1 - i;
        double s = b[i];
        for (int j = nFR - 1; j > d; j--) s -= q->T[(size_t)i * nV + j] * v[j];
        { double piv = q->T[(size_t)i * nV + d]; v[d] = quot(s, piv, 1.0 / piv); } /* pivot reciprocals are formed ahead of the substitution */
    }
}
/* T' u = r, r indexed by Q column, u by AC position */
static void solve_Tt(const orc_qp* q, const double* r, double* u) {
    int nV = q->nV, nFR = q->nFR, nAC = q->nAC;
    for (int i = nAC - 1; i >= 0; i--) {
        int d = nFR - 1 - i;
        double s = r[d];
        for (int k = nAC - 1; k > i; k--) s -= q->T[(size_t)k * nV + d] * u[k];
        { double piv = q->T[(size_t)i * nV + d]; u[i] = quot(s, piv, 1.0 / piv); }
    }
}

/* ------------------------------------------------------------ step direction */
/* Solves the KKT system of the current working set for the data shift
 * (dgv, bound shifts dbF on fixed variables, constraint shifts dbA on active constraints).
 * Outputs q->dx (nV), q->dy (nV+nC).  (qpOASES determineStepDirection) */
static void step_direction_core(orc_qp* q, const double* dgv, const double* dxFX_full, const double* dbAC) {
    int nV = q->nV, nC = q->nC, nFR = q->nFR, nAC = q->nAC, nZ = nFR - nAC;
    double *dx = q->dx, *dy = q->dy;
    for (int i = 0; i < nV; i++) dx[i] = (q->sB[i] != 0) ? dxFX_full[i] : 0.0;
    /* Y part */
    if (nAC > 0) {
        mulA(q, dx, q->t2);
        for (int i = 0; i < nAC; i++) q->t3[i] = dbAC[i] - q->t2[q->AC[i]];
        solve_T(q, q->t3, q->yv);
        for (int p = 0; p < nFR; p++) {
            double s = 0.0;
            for (int j = nZ; j < nFR; j++) s += q->Q[(size_t)p * nV + j] * q->yv[j];
            dx[q->FR[p]] = s;
        }
    }
    /* Z part */
    if (nZ > 0) {
        mulH(q, dx, q->t1);
        for (int j = 0; j < nZ; j++) {
            double s = 0.0;
            for (int p = 0; p < nFR; p++) s += q->Q[(size_t)p * nV + j] * (q->t1[q->FR[p]] + dgv[q->FR[p]]);
            q->zv[j] = -s;
        }
        /* R' u = rhs ; R z = u */
        for (int i = 0; i < nZ; i++) {
            double s = q->zv[i];
            for (int k = 0; k < i; k++) s -= q->R[(size_t)k * nV + i] * q->zv[k];
            { double piv = q->R[(size_t)i * nV + i]; q->zv[i] = quot(s, piv, 1.0 / piv); }
        }
        for (int i = nZ - 1; i >= 0; i--) {
            double s = q->zv[i];
            for (int k = nZ - 1; k > i; k--) s -= q->R[(size_t)i * nV + k] * q->zv[k];
            { double piv = q->R[(size_t)i * nV + i]; q->zv[i] = quot(s, piv, 1.0 / piv); }
        }
        for (int p = 0; p < nFR; p++) {
            double s = 0.0;
            for (int j = 0; j < nZ; j++) s += q->Q[(size_t)p * nV + j] * q->zv[j];
            dx[q->FR[p]] += s;
        }
    }
    /* duals */
    mulH(q, dx, q->t1);
    for (int i = 0; i < nV; i++) q->t1[i] += dgv[i];
    for (int i = 0; i < nV + nC; i++) dy[i] = 0.0;
    if (nAC > 0) {
        for (int j = nZ; j < nFR; j++) {
            double s = 0.0;
            for (int p = 0; p < nFR; p++) s += q->Q[(size_t)p * nV + j] * q->t1[q->FR[p]];
            q->yv[j] = s;
        }
        solve_Tt(q, q->yv, q->t3);
        for (int i = 0; i < nAC; i++) dy[nV + q->AC[i]] = q->t3[i];
        mulAT(q, dy + nV, q->t2);
        for (int i = 0; i < nV; i++) if (q->sB[i] != 0) dy[i] = q->t1[i] - q->t2[i];
    } else {
        for (int i = 0; i < nV; i++) if (q->sB[i] != 0) dy[i] = q->t1[i];
    }
    q->flops += 4.0 * nFR * nFR + 2.0 * nZ * nZ + 2.0 * nAC * nAC;
}

/* numRefinementSteps rounds of iterative refinement (deviation study only: g_refine = 0 in the shipped configuration).
 * The KKT system is linear in its data, so the correction is the same solve applied to the residuals. */
static void step_direction(orc_qp* q, const double* dgv, const double* dxFX_full, const double* dbAC) {
    step_direction_core(q, dgv, dxFX_full, dbAC);
    if (g_refine <= 0) return;
    int nV = q->nV, nC = q->nC, nAC = q->nAC;
    double* dx0 = (double*)malloc(sizeof(double) * (size_t)(nV + 1));
    double* dy0 = (double*)malloc(sizeof(double) * (size_t)(nV + nC + 1));
    double* rg = (double*)malloc(sizeof(double) * (size_t)(nV + 1));
    double* rA = (double*)malloc(sizeof(double) * (size_t)(nC + 1));
    double* zero = (double*)calloc((size_t)(nV + 1), sizeof(double));
    for (int r = 0; r < g_refine; r++) {
        memcpy(dx0, q->dx, sizeof(double) * (size_t)nV);
        memcpy(dy0, q->dy, sizeof(double) * (size_t)(nV + nC));
        mulH(q, dx0, q->t1);
        mulAT(q, dy0 + nV, q->t2);
        mulA(q, dx0, q->t3);
        double rn = 0.0;
        for (int i = 0; i < nV; i++) { rg[i] = (q->sB[i] == 0) ? (q->t1[i] + dgv[i] - q->t2[i]) : 0.0; if (fabs(rg[i]) > rn) rn = fabs(rg[i]); }
        for (int i = 0; i < nAC; i++) { rA[i] = dbAC[i] - q->t3[q->AC[i]]; if (fabs(rA[i]) > rn) rn = fabs(rA[i]); }
        if (rn < 1.0e2 * QP_EPS) break; /* epsIterRef */
        step_direction_core(q, rg, zero, rA);
        for (int i = 0; i < nV; i++) q->dx[i] = dx0[i] + ((q->sB[i] == 0) ? q->dx[i] : 0.0);
        for (int i = 0; i < nC; i++) q->dy[nV + i] += dy0[nV + i];
        /* bound multipliers from stationarity with the refined step */
        mulH(q, q->dx, q->t1);
        mulAT(q, q->dy + nV, q->t2);
        for (int i = 0; i < nV; i++) q->dy[i] = (q->sB[i] != 0) ? (q->t1[i] + dgv[i] - q->t2[i]) : 0.0;
    }
    free(dx0); free(dy0); free(rg); free(rA); free(zero);
}

/* ------------------------------------------------------------ drift correction */
/* (qpOASES performDriftCorrection + setupAuxiliaryQPgradient) */
static void drift_correction(orc_qp* q) {
    int nV = q->nV, nC = q->nC;
    mulA(q, q->x, q->Ax); /* Ax is re-evaluated, not carried along, so it cannot drift from A*x */
    for (int i = 0; i < nV; i++) {
        if (q->sB[i] < 0) { q->lb[i] = q->x[i]; if (q->ub[i] < q->x[i]) q->ub[i] = q->x[i]; if (q->y[i] < 0) q->y[i] = 0.0; }
        else if (q->sB[i] > 0) { q->ub[i] = q->x[i]; if (q->lb[i] > q->x[i]) q->lb[i] = q->x[i]; if (q->y[i] > 0) q->y[i] = 0.0; }
        else { if (q->lb[i] > q->x[i]) q->lb[i] = q->x[i]; if (q->ub[i] < q->x[i]) q->ub[i] = q->x[i]; q->y[i] = 0.0; }
    }
    for (int i = 0; i < nC; i++) {
        double ax = q->Ax[i];
        if (q->sC[i] < 0) { q->lbA[i] = ax; if (q->ubA[i] < ax) q->ubA[i] = ax; if (q->y[nV + i] < 0) q->y[nV + i] = 0.0; }
        else if (q->sC[i] > 0) { q->ubA[i] = ax; if (q->lbA[i] > ax) q->lbA[i] = ax; if (q->y[nV + i] > 0) q->y[nV + i] = 0.0; }
        else { if (q->lbA[i] > ax) q->lbA[i] = ax; if (q->ubA[i] < ax) q->ubA[i] = ax; q->y[nV + i] = 0.0; }
    }
    /* g = A'y_c + y_b - Hx */
    mulAT(q, q->y + nV, q->t2);
    mulH(q, q->x, q->t1);
    for (int i = 0; i < nV; i++) q->g[i] = q->t2[i] + q->y[i] - q->t1[i];
}

/* ------------------------------------------------------------ ramping */
/* After a zero-length homotopy step the current QP data are re-centred on (x, Ax) with a
 * strictly complementary "ramp" of primal slacks and dual values so that ties cannot cycle
 * (qpOASES performRamping: initialRamping = 0.5, finalRamping = 1.0, offset advanced per call). */
static void ramping(orc_qp* q) {
    int nV = q->nV, nC = q->nC;
    int nRamp = nV + nC + nC + nV;
    double r0 = 0.5, r1 = 1.0;
    mulA(q, q->x, q->Ax);
    for (int i = 0; i < nV; i++) {
        double tP = (double)((i + q->ramp_offset) % nRamp) / (double)(nRamp - 1);
        double rP = (1.0 - tP) * r0 + tP * r1;
        double tD = (double)((nV + nC + i + q->ramp_offset) % nRamp) / (double)(nRamp - 1);
        double rD = (1.0 - tD) * r0 + tD * r1;
        double sca = fabs(q->x[i]) > 1.0 ? fabs(q->x[i]) : 1.0;
        if (q->sB[i] >= 0) q->lb[i] = q->x[i] - sca * rP;
        if (q->sB[i] <= 0) q->ub[i] = q->x[i] + sca * rP;
        if (q->sB[i] < 0) { q->lb[i] = q->x[i]; q->y[i] = rD; }
        if (q->sB[i] > 0) { q->ub[i] = q->x[i]; q->y[i] = -rD; }
        if (q->sB[i] == 0) q->y[i] = 0.0;
    }
    for (int i = 0; i < nC; i++) {
        double tP = (double)((nV + i + q->ramp_offset) % nRamp) / (double)(nRamp - 1);
        double rP = (1.0 - tP) * r0 + tP * r1;
        double tD = (double)((nV + nC + nV + i + q->ramp_offset) % nRamp) / (double)(nRamp - 1);
        double rD = (1.0 - tD) * r0 + tD * r1;
        double ax = q->Ax[i];
        double sca = fabs(ax) > 1.0 ? fabs(ax) : 1.0;
        if (q->sC[i] >= 0) q->lbA[i] = ax - sca * rP;
        if (q->sC[i] <= 0) q->ubA[i] = ax + sca * rP;
        if (q->sC[i] < 0) { q->lbA[i] = ax; q->y[nV + i] = rD; }
        if (q->sC[i] > 0) { q->ubA[i] = ax; q->y[nV + i] = -rD; }
        if (q->sC[i] == 0) q->y[nV + i] = 0.0;
    }
    mulAT(q, q->y + nV, q->t2);
    mulH(q, q->x, q->t1);
    for (int i = 0; i < nV; i++) q->g[i] = q->t2[i] + q->y[i] - q->t1[i];
    q->ramp_offset++;
}

/* ------------------------------------------------------------ exchange (ensure LI) */
/* The element to add (constraint c, or bound v when c<0) is linearly dependent on the working
 * set.  Find the combination, run the dual ratio test, remove the blocking element.
 * Returns 0 ok, 1 infeasible. (qpOASES addConstraint_ensureLI / addBound_ensureLI) */
static int ensure_li(orc_qp* q, int c, int v, int status) {
    int nV = q->nV, nC = q->nC, nFR = q->nFR, nAC = q->nAC, nZ = nFR - nAC;
    /* w (over Q columns) is already in q->w; solve T' xiC = w_Y */
    for (int j = nZ; j < nFR; j++) q->yv[j] = q->w[j];
    solve_Tt(q, q->yv, q->xiC);
    for (int i = 0; i < nC; i++) q->t3[i] = 0.0;
    for (int i = 0; i < nAC; i++) q->t3[q->AC[i]] = q->xiC[i];
    mulAT(q, q->t3, q->t2);
    for (int i = 0; i < nV; i++) {
        if (q->sB[i] == 0) { q->xiB[i] = 0.0; continue; }
        double ai = (c >= 0) ? q->Ad[(size_t)c * nV + i] : (i == v ? 1.0 : 0.0);
        q->xiB[i] = ai - q->t2[i];
    }
    double sgn = (status < 0) ? 1.0 : -1.0; /* lower: mu >= 0 ; upper: mu <= 0 */
    double ymin = (g_variant & VAR_MAX_DUAL_JUMP) ? QP_MAX_DUAL_JUMP : QP_INFTY; int kind = -1, idx = -1; /* shipped: no maxDualJump cap (multipliers scale with rho) */
    for (int i = 0; i < nAC; i++) {
        int ci = q->AC[i]; double xi = sgn * q->xiC[i], yy = q->y[nV + ci];
        if (q->sC[ci] < 0) { if (xi > QP_ZERO && yy >= 0.0 && yy / xi < ymin) { ymin = yy / xi; kind = 0; idx = ci; } }
        else { if (xi < -QP_ZERO && yy <= 0.0 && yy / xi < ymin) { ymin = yy / xi; kind = 0; idx = ci; } }
    }
    for (int i = 0; i < nV; i++) {
        if (q->sB[i] == 0) continue;
        double xi = sgn * q->xiB[i], yy = q->y[i];
        if (q->sB[i] < 0) { if (xi > QP_ZERO && yy >= 0.0 && yy / xi < ymin) { ymin = yy / xi; kind = 1; idx = i; } }
        else { if (xi < -QP_ZERO && yy <= 0.0 && yy / xi < ymin) { ymin = yy / xi; kind = 1; idx = i; } }
    }
    if (kind < 0) return 1;
    /* dual update */
    for (int i = 0; i < nAC; i++) q->y[nV + q->AC[i]] -= ymin * sgn * q->xiC[i];
    for (int i = 0; i < nV; i++) if (q->sB[i] != 0) q->y[i] -= ymin * sgn * q->xiB[i];
    if (c >= 0) q->y[nV + c] = sgn * ymin; else q->y[v] = sgn * ymin;
    if (q->verbose > 1) printf("    exchange: remove %s %d (ymin=%g)\n", kind ? "bound" : "constr", idx, ymin);
    if (kind == 0) { q->y[nV + idx] = 0.0; remove_constraint(q, idx); }
    else { q->y[idx] = 0.0; remove_bound(q, idx); }
    /* nZ grew by one; it shrinks again when the new element is added, after which R is recomputed */
    return 0;
}

/* ------------------------------------------------------------ homotopy */
static int homotopy(orc_qp* q, const orc_qp_options* opt) {
    int nV = q->nV, nC = q->nC;
    q->iters = 0;
    for (int it = 0;; it++) {
        /* data shift */
        for (int i = 0; i < nV; i++) { q->dg[i] = q->gN[i] - q->g[i]; q->dlb[i] = q->lbN[i] - q->lb[i]; q->dub[i] = q->ubN[i] - q->ub[i]; }
        for (int i = 0; i < nC; i++) { q->dlbA[i] = q->lbAN[i] - q->lbA[i]; q->dubA[i] = q->ubAN[i] - q->ubA[i]; }
        for (int i = 0; i < nV; i++) q->w[i] = q->sB[i] < 0 ? q->dlb[i] : (q->sB[i] > 0 ? q->dub[i] : 0.0);
        for (int i = 0; i < q->nAC; i++) { int ci = q->AC[i]; q->a[i] = q->sC[ci] < 0 ? q->dlbA[ci] : q->dubA[ci]; }
        step_direction(q, q->dg, q->w, q->a);
        mulA(q, q->dx, q->dAx);

        /* ratio tests (qpOASES performStep / performRatioTest / isBlocking) */
        double tau = 1.0; int bc_idx = -1, bc_isbound = 0, bc_status = 0;
/* isBlocking: den >= epsDen, num >= epsNum, num < den (the full step is not reachable); the
         * ratio then has to beat the running minimum strictly, so the first index scanned wins ties */
#define BLOCKING(num, den) ((den) >= QP_EPS_DEN && (num) >= QP_EPS_NUM && ((g_variant & VAR_RATIO_MULT) ? ((num) < tau * (den)) : ((num) < (den) && (num) / (den) < tau)))
        for (int i = 0; i < q->nAC; i++) { /* duals of active constraints */
            int ci = q->AC[i]; double num, den;
            if (q->sC[ci] < 0) { num = q->y[nV + ci]; den = -q->dy[nV + ci]; } else { num = -q->y[nV + ci]; den = q->dy[nV + ci]; }
            if (BLOCKING(num, den)) { tau = num / den; bc_idx = ci; bc_isbound = 0; bc_status = 0; }
        }
        for (int i = 0; i < nV; i++) { /* duals of fixed bounds */
            if (q->sB[i] == 0) continue; double num, den;
            if (q->sB[i] < 0) { num = q->y[i]; den = -q->dy[i]; } else { num = -q->y[i]; den = q->dy[i]; }
            if (BLOCKING(num, den)) { tau = num / den; bc_idx = i; bc_isbound = 1; bc_status = 0; }
        }
        for (int i = 0; i < nC; i++) { /* inactive constraints, lower */
            if (q->sC[i] != 0) continue;
            double num = q->Ax[i] - q->lbA[i]; if (num < 0) num = 0; double den = q->dlbA[i] - q->dAx[i];
            if (BLOCKING(num, den)) { tau = num / den; bc_idx = i; bc_isbound = 0; bc_status = -1; }
        }
        for (int i = 0; i < nC; i++) { /* inactive constraints, upper */
            if (q->sC[i] != 0) continue;
            double num = q->ubA[i] - q->Ax[i]; if (num < 0) num = 0; double den = q->dAx[i] - q->dubA[i];
            if (BLOCKING(num, den)) { tau = num / den; bc_idx = i; bc_isbound = 0; bc_status = 1; }
        }
        for (int i = 0; i < nV; i++) { /* free variables, lower */
            if (q->sB[i] != 0) continue;
            double num = q->x[i] - q->lb[i]; if (num < 0) num = 0; double den = q->dlb[i] - q->dx[i];
            if (BLOCKING(num, den)) { tau = num / den; bc_idx = i; bc_isbound = 1; bc_status = -1; }
        }
        for (int i = 0; i < nV; i++) { /* free variables, upper */
            if (q->sB[i] != 0) continue;
            double num = q->ub[i] - q->x[i]; if (num < 0) num = 0; double den = q->dx[i] - q->dub[i];
            if (BLOCKING(num, den)) { tau = num / den; bc_idx = i; bc_isbound = 1; bc_status = 1; }
        }
#undef BLOCKING
        if (q->verbose) printf("  it %d: tau=%.6e  bc=%s%d -> %d  nFR=%d nAC=%d\n", it, tau,
                               bc_idx < 0 ? "none" : (bc_isbound ? "b" : "c"), bc_idx, bc_status, q->nFR, q->nAC);
        if (q->verbose > 2) {
            printf("    FR:"); for (int p = 0; p < q->nFR; p++) printf(" %d", q->FR[p]);
            printf("  AC:"); for (int p = 0; p < q->nAC; p++) printf(" %d", q->AC[p]);
            printf("\n    x :"); for (int i = 0; i < nV; i++) printf(" %.17g", q->x[i]);
            printf("\n    dx:"); for (int i = 0; i < nV; i++) printf(" %.17g", q->dx[i]);
            printf("\n    Ax:"); for (int i = 0; i < nC; i++) printf(" %.17g", q->Ax[i]);
            printf("\n    dAx:"); for (int i = 0; i < nC; i++) printf(" %.17g", q->dAx[i]);
            printf("\n    ubA:"); for (int i = 0; i < nC; i++) printf(" %.17g", q->ubA[i]);
            printf("\n    T:"); for (int i = 0; i < q->nAC; i++) for (int j = 0; j < q->nFR; j++) printf(" %.17g", q->T[(size_t)i*nV+j]);
            printf("\n    Q:"); for (int i = 0; i < q->nFR; i++) for (int j = 0; j < q->nFR; j++) printf(" %.17g", q->Q[(size_t)i*nV+j]);
            printf("\n");
        }
        /* step */
        if (bc_idx < 0) {
            for (int i = 0; i < nV; i++) { q->x[i] += q->dx[i]; q->g[i] = q->gN[i]; q->lb[i] = q->lbN[i]; q->ub[i] = q->ubN[i]; }
            for (int i = 0; i < nV + nC; i++) q->y[i] += q->dy[i];
            for (int i = 0; i < nC; i++) { q->Ax[i] += q->dAx[i]; q->lbA[i] = q->lbAN[i]; q->ubA[i] = q->ubAN[i]; }
            q->iters = it;
            return ORC_QP_OPTIMAL;
        }
        if (it >= opt->max_iter) { q->iters = it; return ORC_QPERROR_PERFORMINGHOMOTOPY; }
        if (tau > 0.0) {
            for (int i = 0; i < nV; i++) { q->x[i] += tau * q->dx[i]; q->g[i] += tau * q->dg[i]; q->lb[i] += tau * q->dlb[i]; q->ub[i] += tau * q->dub[i]; }
            for (int i = 0; i < nV + nC; i++) q->y[i] += tau * q->dy[i];
            for (int i = 0; i < nC; i++) { q->Ax[i] += tau * q->dAx[i]; q->lbA[i] += tau * q->dlbA[i]; q->ubA[i] += tau * q->dubA[i]; }
        }
        /* change the working set */
        if (bc_status == 0) {
            int flipped = 0;
            if (bc_isbound) {
                int old = q->sB[bc_idx];
                q->y[bc_idx] = 0.0;
                remove_bound(q, bc_idx);
                if (!extend_R(q, opt->enable_flipping)) {
                    /* no positive curvature: flip to the opposite bound */
                    if (!opt->enable_flipping) return ORC_QPERROR_UNBOUNDED;
                    bound_w(q, bc_idx);
                    add_bound(q, bc_idx, -old);
                    if (old < 0) q->ub[bc_idx] = q->x[bc_idx]; else q->lb[bc_idx] = q->x[bc_idx];
                    flipped = 1;
                }
            } else {
                int old = q->sC[bc_idx];
                q->y[nV + bc_idx] = 0.0;
                remove_constraint(q, bc_idx);
                if (!extend_R(q, opt->enable_flipping)) {
                    if (!opt->enable_flipping) return ORC_QPERROR_UNBOUNDED;
                    double z2, a2;
                    constraint_w(q, bc_idx, &z2, &a2);
                    add_constraint(q, bc_idx, -old);
                    if (old < 0) q->ubA[bc_idx] = q->Ax[bc_idx]; else q->lbA[bc_idx] = q->Ax[bc_idx];
                    flipped = 1;
                }
            }
            if (flipped) {
                if (q->verbose) printf("    flipped %s %d\n", bc_isbound ? "bound" : "constr", bc_idx);
                if (recompute_R(q)) { q->iters = it; return ORC_QPERROR_INTERNAL_ERROR; }
            }
        } else {
            if (bc_isbound) {
                double z2 = bound_w(q, bc_idx);
                if (!(z2 > QP_EPS_LI * QP_EPS_LI)) {
                    if (ensure_li(q, -1, bc_idx, bc_status)) { q->iters = it; return ORC_QPERROR_INFEASIBLE; }
                    bound_w(q, bc_idx);
                }
                add_bound(q, bc_idx, bc_status);
            } else {
                double z2, a2;
                constraint_w(q, bc_idx, &z2, &a2);
                if (!(z2 > QP_EPS_LI * QP_EPS_LI * a2) || a2 == 0.0) {
                    if (ensure_li(q, bc_idx, -1, bc_status)) { q->iters = it; return ORC_QPERROR_INFEASIBLE; }
                    constraint_w(q, bc_idx, &z2, &a2);
                }
                add_constraint(q, bc_idx, bc_status);
            }
            if (recompute_R(q)) { q->iters = it; return ORC_QPERROR_INTERNAL_ERROR; }
        }
        /* zero step: ramping; otherwise drift correction (qpOASES solveQP, step 4) */
        if (tau <= QP_EPS && opt->enable_ramping) ramping(q);
        else if (opt->enable_drift) drift_correction(q);
    }
}

/* ------------------------------------------------------------ public entry points */
static double clampinf(double v) { return v > QP_INFTY ? QP_INFTY : (v < -QP_INFTY ? -QP_INFTY : v); }

static void set_targets(orc_qp* q, const double* g, const double* lb, const double* ub,
                        const double* lbA, const double* ubA) {
    for (int i = 0; i < q->nV; i++) { q->gN[i] = g[i]; q->lbN[i] = clampinf(lb[i]); q->ubN[i] = clampinf(ub[i]); }
    for (int i = 0; i < q->nC; i++) { q->lbAN[i] = clampinf(lbA[i]); q->ubAN[i] = clampinf(ubA[i]); }
}

static void copy_csc(int ncol, const int* p, const int* i, const double* v, int** P, int** I, double** V) {
    int nnz = p[ncol];
    free(*P); free(*I); free(*V);
    *P = (int*)zalloc(ncol + 1, 4); *I = (int*)zalloc(nnz, 4); *V = (double*)zalloc(nnz, 8);
    memcpy(*P, p, sizeof(int) * (ncol + 1));
    memcpy(*I, i, sizeof(int) * nnz);
    memcpy(*V, v, sizeof(double) * nnz);
}

/* auxiliary QP of the cold start: x = 0, y = 0, all bounds active at lower, no constraint active */
static int cold_start(orc_qp* q, const orc_qp_options* opt) {
    int nV = q->nV, nC = q->nC;
    q->nFR = 0; q->nAC = 0; q->ramp_offset = 0;
    for (int i = 0; i < nV; i++) {
        q->x[i] = 0.0; q->y[i] = 0.0; q->sB[i] = -1; q->posFR[i] = -1;
        q->g[i] = 0.0; q->lb[i] = 0.0; q->ub[i] = QP_BOUND_RELAX;
    }
    for (int i = 0; i < nC; i++) {
        q->y[nV + i] = 0.0; q->sC[i] = 0; q->posAC[i] = -1; q->Ax[i] = 0.0;
        q->lbA[i] = -QP_BOUND_RELAX; q->ubA[i] = QP_BOUND_RELAX;
    }
    q->last_cold = 1;
    q->status = homotopy(q, opt);
    q->initialised = 1;
    return q->status;
}

/* Auxiliary QP of an init from a guess (qpOASES QProblem::solveInitialQP as reached from src/qpOASESInterface.cpp:202-207 and
 * :716-729): x, y, A x and the wanted bound statuses sB are set by the caller, want[i] is the wanted status of constraint i.
 * Working set: bounds first, then the constraints in index order, linearly dependent ones left out; auxiliary bounds equal to
 * x / A x on the active side and relaxed by boundRelaxation elsewhere -- except that a constraint left out for dependence
 * keeps its wanted side tight ("strongly inactive"); gradient from stationarity; projected Cholesky factor. */
static int aux_qp_from_guess(orc_qp* q, const int* want) {
    int nV = q->nV, nC = q->nC;
    q->nFR = 0; q->nAC = 0; q->ramp_offset = 0;
    for (int i = 0; i < nV; i++) {
        double xi = q->x[i];
        int st = q->sB[i];
        q->posFR[i] = -1;
        q->lb[i] = (st < 0) ? xi : xi - QP_BOUND_RELAX;
        q->ub[i] = (st > 0) ? xi : xi + QP_BOUND_RELAX;
    }
    for (int i = 0; i < nC; i++) { q->sC[i] = 0; q->posAC[i] = -1; }
    for (int i = 0; i < nV; i++) if (q->sB[i] == 0) { q->FR[q->nFR] = i; q->posFR[i] = q->nFR; q->nFR++; }
    if (q->nFR > q->max_nFR) q->max_nFR = q->nFR;
    for (int i = 0; i < q->nFR; i++)
        for (int j = 0; j < q->nFR; j++) q->Q[(size_t)i * nV + j] = (i == j) ? 1.0 : 0.0;
    for (int i = 0; i < nC; i++) {
        double ax = q->Ax[i];
        int st = want[i];
        q->lbA[i] = (st < 0) ? ax : ax - QP_BOUND_RELAX;
        q->ubA[i] = (st > 0) ? ax : ax + QP_BOUND_RELAX;
        if (st == 0 || q->nAC >= q->nFR) continue;
        double z2, a2;
        constraint_w(q, i, &z2, &a2);
        if (!(z2 > QP_EPS_LI * QP_EPS_LI * a2) || a2 == 0.0) continue;
        add_constraint(q, i, st);
    }
    mulAT(q, q->y + nV, q->t2);
    mulH(q, q->x, q->t1);
    for (int i = 0; i < nV; i++) q->g[i] = q->t2[i] + q->y[i] - q->t1[i];
    q->initialised = 1;
    return recompute_R(q);
}

/* init(H, g, A, lb, ub, lbA, ubA, nWSR, 0, x0) of handle_error's infeasible branch (src/qpOASESInterface.cpp:716-729, 690-701):
 * primal guess x0 = [0; max(0, lbA); -min(0, ubA)] (the slack-feasible point of the l1-penalty QP), y = 0; working set read
 * off x0 and A x0 with boundTolerance = 1e6*EPS; then the homotopy to the real data. */
static int guess_start(orc_qp* q, const orc_qp_options* opt) {
    int nV = q->nV, nC = q->nC, o1 = nV - 2 * nC, o2 = nV - nC;
    const double TOL = 1.0e6 * QP_EPS;
    int* want = (int*)zalloc(nC ? nC : 1, 4);
    q->last_cold = 1;
    for (int i = 0; i < nV; i++) { q->x[i] = 0.0; q->y[i] = 0.0; }
    for (int i = 0; i < nC; i++) { q->x[o1 + i] = fmax(0.0, q->lbAN[i]); q->x[o2 + i] = -fmin(0.0, q->ubAN[i]); q->y[nV + i] = 0.0; }
    mulA(q, q->x, q->Ax);
    for (int i = 0; i < nV; i++) { double xi = q->x[i]; q->sB[i] = (xi <= q->lbN[i] + TOL) ? -1 : ((xi >= q->ubN[i] - TOL) ? 1 : 0); }
    for (int i = 0; i < nC; i++) { double ax = q->Ax[i]; want[i] = (ax <= q->lbAN[i] + TOL) ? -1 : ((ax >= q->ubAN[i] - TOL) ? 1 : 0); }
    int bad = aux_qp_from_guess(q, want);
    free(want);
    if (bad) { q->iters = 0; q->status = ORC_QPERROR_INTERNAL_ERROR; return q->status; }
    q->status = homotopy(q, opt);
    return q->status;
}

/* init(H, g, A, lb, ub, lbA, ubA, nWSR, 0, x_qp, y_qp, &bounds) of a FIXED <-> VARIED flip of the matrix status
 * (src/qpOASESInterface.cpp:202-207): primal and dual guess = the previous solution, bound statuses = the previous ones,
 * constraint statuses from the sign of the guessed multipliers (y > EPS lower, y < -EPS upper; qpOASES
 * obtainAuxiliaryWorkingSet with guessedConstraints == 0 and yOpt != 0), new matrices.  1 if the projected Hessian of that
 * working set is not positive definite (the init fails before its homotopy). */
static int reinit_state(orc_qp* q) {
    int nV = q->nV, nC = q->nC;
    int* want = (int*)zalloc(nC ? nC : 1, 4);
    for (int i = 0; i < nC; i++) { double yi = q->y[nV + i]; want[i] = (yi > QP_EPS) ? -1 : ((yi < -QP_EPS) ? 1 : 0); }
    mulA(q, q->x, q->Ax);
    int bad = aux_qp_from_guess(q, want);
    free(want);
    return bad;
}

/* qpOASESInterface::handle_error (src/qpOASESInterface.cpp:686-758) on top of this solver; the backend restatements call it
 * whenever a solve did not end OPTIMAL (after a hot start AND after an init, :160-162, :217-219).  Infeasible (or
 * force_guess, a test hook): re-init from the slack-feasible guess.  Otherwise a plain re-init -- which after a failed init
 * repeats the same deterministic solve, so only its iteration count is added again (the reference adds nWSR of both runs to
 * Stats::qp_iter, :752-753).  Returns the status; *iters_added = iterations of the recovery attempt. */
int orc_qp_handle_error(orc_qp* q, const orc_qp_options* opt, int force_guess, int* iters_added) {
    int st;
    if (q->fell_back && !force_guess) { /* the cold start run inside hotstart_matrices already was the recovery attempt */
        if (iters_added) *iters_added = 0;
        return q->status;
    }
    if ((q->status == ORC_QPERROR_INFEASIBLE || force_guess) && q->nV >= 2 * q->nC) st = guess_start(q, opt);
    else if (!q->last_cold) st = cold_start(q, opt);
    else st = q->status;
    if (iters_added) *iters_added = q->iters;
    return st;
}

int orc_qp_init(orc_qp* q, const orc_qp_options* opt, const int* H_colptr, const int* H_rowidx,
                const double* H_val, const double* g, const int* A_colptr, const int* A_rowidx,
                const double* A_val, const double* lb, const double* ub, const double* lbA,
                const double* ubA, int is_lp) {
    int nV = q->nV;
    q->has_H = (H_colptr != NULL) && !is_lp;
    q->is_lp = is_lp || H_colptr == NULL;
    q->reg = q->is_lp ? QP_EPS_REG : 0.0;
    if (q->has_H) copy_csc(nV, H_colptr, H_rowidx, H_val, &q->Hp, &q->Hi, &q->Hv);
    copy_csc(nV, A_colptr, A_rowidx, A_val, &q->Ap, &q->Ai, &q->Av);
    build_dense_A(q);
    set_targets(q, g, lb, ub, lbA, ubA);
    q->flops = 0.0;
    q->fell_back = 0;
    return cold_start(q, opt);
}

int orc_qp_hotstart(orc_qp* q, const orc_qp_options* opt, const double* g, const double* lb,
                    const double* ub, const double* lbA, const double* ubA) {
    if (!q->initialised) return ORC_QPERROR_NOTINITIALISED;
    set_targets(q, g, lb, ub, lbA, ubA);
    q->last_cold = 0; q->fell_back = 0;
    q->status = homotopy(q, opt);
    return q->status;
}

/* Rebuild TQ and R for the kept working set with new matrix values
 * (qpOASES SQProblem::hotstart -> setupAuxiliaryQP). */
static int refactorise(orc_qp* q) {
    int nV = q->nV, nFR = q->nFR;
    int nAC_old = q->nAC;
    int* ac = (int*)zalloc(nAC_old, 4);
    int* st = (int*)zalloc(nAC_old, 4);
    for (int i = 0; i < nAC_old; i++) { ac[i] = q->AC[i]; st[i] = q->sC[ac[i]]; }
    for (int i = 0; i < nFR; i++)
        for (int j = 0; j < nFR; j++) q->Q[(size_t)i * nV + j] = (i == j) ? 1.0 : 0.0;
    q->nAC = 0;
    for (int i = 0; i < nAC_old; i++) { q->sC[ac[i]] = 0; q->posAC[ac[i]] = -1; }
    for (int i = 0; i < nAC_old; i++) {
        double z2, a2;
        constraint_w(q, ac[i], &z2, &a2);
        if (!(z2 > QP_EPS_LI * QP_EPS_LI * a2) || a2 == 0.0) { q->y[nV + ac[i]] = 0.0; continue; } /* dependent: drop */
        add_constraint(q, ac[i], st[i]);
    }
    free(ac); free(st);
    return recompute_R(q);
}

int orc_qp_hotstart_matrices(orc_qp* q, const orc_qp_options* opt, const double* H_val,
                             const double* A_val, const double* g, const double* lb,
                             const double* ub, const double* lbA, const double* ubA) {
    if (!q->initialised) return ORC_QPERROR_NOTINITIALISED;
    int nV = q->nV;
    if (q->has_H && H_val) memcpy(q->Hv, H_val, sizeof(double) * q->Hp[nV]);
    if (A_val) { memcpy(q->Av, A_val, sizeof(double) * q->Ap[nV]); build_dense_A(q); }
    set_targets(q, g, lb, ub, lbA, ubA);
    q->fell_back = 0;
    if (refactorise(q)) {
        q->fell_back = 1;
        /* projected Hessian of the kept working set is not positive definite: cold start
         * (what the reference does through handle_error, src/qpOASESInterface.cpp:746-749) */
        return cold_start(q, opt);
    }
    mulA(q, q->x, q->Ax);
    drift_correction(q);
    q->last_cold = 0;
    q->status = homotopy(q, opt);
    return q->status;
}

/* the matrix-status flip of optimizeQP (src/qpOASESInterface.cpp:202-207): a fresh init from the previous solution */
int orc_qp_reinit(orc_qp* q, const orc_qp_options* opt, const double* H_val, const double* A_val, const double* g,
                  const double* lb, const double* ub, const double* lbA, const double* ubA) {
    if (!q->initialised) return ORC_QPERROR_NOTINITIALISED;
    int nV = q->nV;
    if (q->has_H && H_val) memcpy(q->Hv, H_val, sizeof(double) * q->Hp[nV]);
    if (A_val) { memcpy(q->Av, A_val, sizeof(double) * q->Ap[nV]); build_dense_A(q); }
    set_targets(q, g, lb, ub, lbA, ubA);
    q->fell_back = 0;
    if (reinit_state(q)) { q->fell_back = 1; return cold_start(q, opt); } /* as hotstart_matrices: the plain re-init of handle_error */
    q->last_cold = 0; /* a failure is followed by handle_error's plain re-init, a different solve */
    q->status = homotopy(q, opt);
    return q->status;
}

void orc_qp_get_solution(const orc_qp* q, double* x, double* y, double* obj, int* iters) {
    int nV = q->nV, nC = q->nC;
    if (x) memcpy(x, q->x, sizeof(double) * nV);
    if (y) memcpy(y, q->y, sizeof(double) * (nV + nC));
    if (obj) { /* 1/2 x'Hx + g'x with the (unregularised) Hessian */
        double s = 0.0;
        if (q->has_H) {
            double* hx = (double*)zalloc(nV, 8);
            for (int c = 0; c < nV; c++)
                for (int e = q->Hp[c]; e < q->Hp[c + 1]; e++) hx[q->Hi[e]] += q->Hv[e] * q->x[c];
            for (int i = 0; i < nV; i++) s += 0.5 * q->x[i] * hx[i];
            free(hx);
        }
        for (int i = 0; i < nV; i++) s += q->gN[i] * q->x[i];
        /* getObjVal() of a problem that is not solved is INFTY (qpOASES QProblemB::getObjVal; src/qpOASESInterface.cpp:324-327) */
        *obj = (q->status == ORC_QP_OPTIMAL) ? s : QP_INFTY;
    }
    if (iters) *iters = q->iters;
}

/* raw convention of getWorkingSetBounds/Constraints as consumed at
 * src/qpOASESInterface.cpp:847-887: +1 upper, -1 lower, 0 inactive */
void orc_qp_get_working_set(const orc_qp* q, int* raw_b, int* raw_c) {
    for (int i = 0; i < q->nV; i++) raw_b[i] = q->sB[i];
    for (int i = 0; i < q->nC; i++) raw_c[i] = q->sC[i];
}

double orc_qp_get_flops(const orc_qp* q) { return q->flops; }
/* largest number of free variables seen since creation (sizes the factors of the CUDA kernel) */
int orc_qp_get_max_free(const orc_qp* q) { return q->max_nFR; }
int orc_qp_get_fell_back(const orc_qp* q) { return q->fell_back; }
