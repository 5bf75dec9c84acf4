// qp_kernel.cuh -- batched FP64 active-set QP/LP solver for sm_100a (row D of SURVEY.md 8a).
//
// One QP per *team* of TEAM threads (32 = one warp, 64..256 = several warps joined by a named
// barrier); a CTA hosts CTA_THREADS/TEAM teams.  All factors (Q of the TQ factorisation; T and
// the projected Cholesky factor R packed into one array) and all iterate/work vectors of a QP
// live in that team's slice of shared memory for the whole solve; HBM is touched only for the
// compulsory input (matrix values, g, bounds) and output (x, y, working set, status, KKT
// residuals) and, when hot starts are requested, for one save/restore of the slice.
//
// The algorithm is the online active-set strategy as the reference drives it through qpOASES
// (call sites src/qpOASESInterface.cpp:155-206, 231-268, 221-222, 843-844, options :765): same
// steps, thresholds and tie-breaks as the CPU oracle (oracle/oracle_qp.c), which is only a
// checker and is never called from here.
//
// Parallelisation inside a team (L = TEAM lanes):
//   * sparse products: one lane per output entry (CSC columns for H and A', a CSR view for A),
//     accumulation in storage order (= the reference's SpHbMat::times order);
//   * Givens sweeps: the rotation chain is a scalar recurrence; every lane applies the whole
//     chain to its own rows of Q / T, so the sweep needs no barrier per rotation when the
//     chain is known up front (additions), and one barrier per rotation otherwise (removals);
//   * triangular solves: column-oriented substitution, lanes over the trailing vector;
//   * projected Hessian + Cholesky: one sparse product per null-space column, row-wise Cholesky
//     with lanes over the columns of the current row;
//   * ratio tests: lane-local scan in index order + (ratio, position) lexicographic min reduction,
//     which reproduces the sequential "first index wins ties" rule.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqpb200 {

// Out-of-line member functions keep the one-warp-per-QP kernel at ~14k SASS instructions (vs ~52k fully
// inlined).  The multi-warp variants (TEAM >= 64) MUST be compiled with -DQP_INLINE_ALL: CTA-level barriers
// (bar.sync / __syncthreads) inside out-of-line functions were observed on sm_100a (CUDA 12.9) to lose
// synchronisation when a warp reaches the call lane-divergent (intermittent wrong results / deadlocks with
// __noinline__, none in 384-instance stress runs when inlined); warp barriers (__syncwarp) are unaffected.
#ifdef QP_INLINE_ALL
#define QP_FN __forceinline__
#else
#define QP_FN __noinline__
#endif

#define QP_EPS 2.221e-16
#define QP_INFTY 1.0e20
#define QP_ZERO 1.0e-25
#define QP_BOUND_RELAX 1.0e4
#define QP_EPS_NUM (-1.0e3 * QP_EPS)
#define QP_EPS_DEN (1.0e3 * QP_EPS)
#define QP_EPS_FLIP (1.0e3 * QP_EPS)
#define QP_EPS_REG (1.0e3 * QP_EPS)
#define QP_EPS_LI (1.0e5 * QP_EPS)

enum { ST_OPTIMAL = 20, ST_INTERNAL = 21, ST_INFEASIBLE = 22, ST_UNBOUNDED = 23, ST_NOTINIT = 25, ST_HOMOTOPY = 28 };
enum { MODE_COLD = 0, MODE_HOT_FIXED = 1, MODE_HOT_VARIED = 2 };
enum { FLAG_FLIPPING = 1, FLAG_RAMPING = 2, FLAG_DRIFT = 4, FLAG_KEEP_STATE = 8 };

struct QPKernelArgs {
    int batch, nV, nC, ld;
    int is_lp, has_H;
    int max_iter, flags, mode;
    int zA, zH;
    // shared sparsity pattern
    const int *Ap, *Ai;            // CSC of A (nC x nV)
    const int *Arp, *Aci, *Aperm;  // CSR view of A: rowptr, column index, position in the CSC value array
    const int *Hp, *Hi;            // CSC of H (nV x nV, full symmetric)
    // per-instance inputs, instance-major
    const double *Aval, *Hval;
    const double *gN, *lbN, *ubN, *lbAN, *ubAN;
    const unsigned char* mask;
    const int* inst_mode;  // optional per-instance mode override (NULL: args.mode)
    // outputs
    double *x, *y, *obj, *kkt;
    int *status, *iters;
    signed char *wsB, *wsC;  // raw working set (+1 upper, -1 lower, 0 inactive)
    int *WB, *WC;            // translated ActiveType
    // resident hot-start state
    double* state;           // [batch][slice_doubles]
    int* state_hdr;          // [batch][4]: nFR, nAC, ramp_offset, initialised
    int slice_doubles;       // doubles per team slice (double part + int part rounded up)
};

__host__ __device__ inline int qp_slice_doubles(int nV, int nC, int ld, int zA, int zH) {
    int nT = nV + nC;
    int d = 2 * nV * ld + 5 * nV + 4 * nC + 9 * nT + zA + zH;
    int shorts = 3 * nV + 3 * nC;
    return d + (shorts * 2 + 7) / 8 + 1;
}

template <int TEAM>
struct Team {
    int lane;     // 0..TEAM-1
    int team_id;  // team index within the CTA
    // TEAM == 32: warp barrier.  TEAM >= 128: the team is the whole CTA (one QP per CTA), so the compiler-known
    // __syncthreads() is used.  TEAM == 64: two warps joined by a named barrier; the warp is re-converged first
    // and the non-.aligned barrier.sync form is used, so a lane-divergent prologue cannot make it undefined.
    __device__ __forceinline__ void sync() const {
        if (TEAM == 32) __syncwarp();
        else if (TEAM >= 128) __syncthreads();
        else {
            __syncwarp();
            asm volatile("barrier.sync %0, %1;" ::"r"(team_id + 1), "r"(TEAM) : "memory");
        }
    }
};

struct MinKey {
    double t;
    int pos;
};
__device__ __forceinline__ bool key_less(double t1, int p1, double t2, int p2) { return t1 < t2 || (t1 == t2 && p1 < p2); }

template <int TEAM>
struct QPSolver {
    Team<TEAM> tm;
    int lane;
    int nV, nC, nT, ld;
    int nFR, nAC, ramp_offset;
    int is_lp, has_H, flags;
    double reg;
    // shared memory
    double *Q, *RT, *x, *y, *Ax, *g, *lb, *ub, *lbA, *ubA, *dx, *dy, *dAx, *t1, *t2, *t3, *w, *a, *yv, *zv, *Av, *Hv;
    short *sB, *sC, *FR, *AC, *posFR, *posAC;
    double* red;  // reduction scratch, per CTA team (TEAM/32 * 2 doubles)
    // global
    const int *Ap, *Ai, *Arp, *Aci, *Aperm, *Hp, *Hi;
    const double *gN, *lbN, *ubN, *lbAN, *ubAN;
    int iters;

    __device__ __forceinline__ void sync() { tm.sync(); }
    __device__ __forceinline__ double& R(int a_, int b_) { return RT[a_ * ld + b_]; }
    __device__ __forceinline__ double& T(int i, int j) { return RT[(nV - 1 - i) * ld + j]; }

    __device__ void carve(double* base, const QPKernelArgs& A) {
        double* p = base;
        Q = p; p += nV * ld;
        RT = p; p += nV * ld;
        x = p; p += nV; g = p; p += nV; lb = p; p += nV; ub = p; p += nV; dx = p; p += nV;
        Ax = p; p += nC; lbA = p; p += nC; ubA = p; p += nC; dAx = p; p += nC;
        y = p; p += nT; dy = p; p += nT; t1 = p; p += nT; t2 = p; p += nT; t3 = p; p += nT;
        w = p; p += nT; a = p; p += nT; yv = p; p += nT; zv = p; p += nT;
        Av = p; p += A.zA; Hv = p; p += A.zH;
        short* s = reinterpret_cast<short*>(p);
        sB = s; s += nV; FR = s; s += nV; posFR = s; s += nV;
        sC = s; s += nC; AC = s; s += nC; posAC = s; s += nC;
    }

    // ---------------------------------------------------------------- sparse products
    __device__ QP_FN void mulH(const double* v, double* out) {  // out = (H + reg I) v   (H symmetric: column gather)
        for (int c = lane; c < nV; c += TEAM) {
            double s = 0.0;
            if (has_H) {
                int e1 = Hp[c + 1];
                for (int e = Hp[c]; e < e1; e++) s += Hv[e] * v[Hi[e]];
            }
            if (reg != 0.0) s += reg * v[c];
            out[c] = s;
        }
        sync();
    }
    __device__ QP_FN void mulA(const double* v, double* out) {
        for (int r = lane; r < nC; r += TEAM) {
            double s = 0.0;
            int k1 = Arp[r + 1];
            for (int k = Arp[r]; k < k1; k++) s += Av[Aperm[k]] * v[Aci[k]];
            out[r] = s;
        }
        sync();
    }
    __device__ QP_FN void mulAT(const double* yc, double* out) {
        for (int c = lane; c < nV; c += TEAM) {
            double s = 0.0;
            int e1 = Ap[c + 1];
            for (int e = Ap[c]; e < e1; e++) s += Av[e] * yc[Ai[e]];
            out[c] = s;
        }
        sync();
    }
    __device__ __forceinline__ double A_entry(int r, int c) {  // A[r][c] via the CSR view (duplicates summed)
        double s = 0.0;
        int k1 = Arp[r + 1];
        for (int k = Arp[r]; k < k1; k++)
            if (Aci[k] == c) s += Av[Aperm[k]];
        return s;
    }

    // ---------------------------------------------------------------- reductions
    __device__ QP_FN MinKey team_min(double t, int pos) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double t2_ = __shfl_xor_sync(0xffffffffu, t, o);
            int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
            if (key_less(t2_, p2, t, pos)) { t = t2_; pos = p2; }
        }
        if (TEAM > 32) {
            int wid = lane >> 5;
            sync();
            if ((lane & 31) == 0) { red[2 * wid] = t; red[2 * wid + 1] = (double)pos; }
            sync();
            t = red[0]; pos = (int)red[1];
            for (int k = 1; k < TEAM / 32; k++) {
                double tk = red[2 * k]; int pk = (int)red[2 * k + 1];
                if (key_less(tk, pk, t, pos)) { t = tk; pos = pk; }
            }
            sync();
        }
        MinKey r; r.t = t; r.pos = pos;
        return r;
    }

    // ---------------------------------------------------------------- Givens
    __device__ __forceinline__ void givens(double a_, double b_, double& c, double& s, double& r) {
        if (a_ == 0.0) { c = 1.0; s = 0.0; r = b_; return; }
        double h = sqrt(a_ * a_ + b_ * b_);
        c = b_ / h; s = a_ / h; r = h;
    }

    // ---------------------------------------------------------------- projected Cholesky
    __device__ QP_FN void proj_column(int b) {  // t2 = (H+regI) * (column b of Q scattered to full space)
        for (int i = lane; i < nV; i += TEAM) { int p = posFR[i]; t1[i] = (p >= 0) ? Q[p * ld + b] : 0.0; }
        sync();
        mulH(t1, t2);
    }
    // returns 0 ok, 1+j on failure (uniform)
    __device__ QP_FN int recompute_R() {
        int nZ = nFR - nAC;
        if (nZ <= 0) return 0;
        if (is_lp) {
            double sr = sqrt(reg);
            for (int k = lane; k < nZ * nZ; k += TEAM) { int a_ = k / nZ, b_ = k % nZ; R(a_, b_) = (a_ == b_) ? sr : 0.0; }
            sync();
            return 0;
        }
        for (int b = 0; b < nZ; b++) {
            proj_column(b);
            for (int a_ = lane; a_ <= b; a_ += TEAM) {
                double s = 0.0;
                for (int p = 0; p < nFR; p++) s += Q[p * ld + a_] * t2[FR[p]];
                R(a_, b) = s;
            }
            sync();
        }
        // row-wise Cholesky: R'R = M, same per-element summation order as the column version
        for (int i = 0; i < nZ; i++) {
            // phase A: s_j = M[i][j] - sum_{k<i} R[k][i] R[k][j]  for j >= i
            for (int j = i + lane; j < nZ; j += TEAM) {
                double s = R(i, j);
                for (int k = 0; k < i; k++) s -= R(k, i) * R(k, j);
                R(i, j) = s;
            }
            sync();
            double d = R(i, i);
            sync();  // all lanes hold d before anyone rewrites R(i,i): the branch below is team-uniform
            if (!(d > QP_ZERO)) return 1 + i;
            double dd = sqrt(d);
            for (int j = i + lane; j < nZ; j += TEAM) R(i, j) = (j == i) ? dd : R(i, j) / dd;
            for (int j = lane; j < i; j += TEAM) R(i, j) = 0.0;
            sync();
        }
        return 0;
    }
    // border R with the new last null-space column; returns 1 if curvature acceptable
    __device__ QP_FN int extend_R(int check_curvature) {
        int nZ = nFR - nAC, b = nZ - 1;
        if (is_lp) {
            for (int a_ = lane; a_ < b; a_ += TEAM) { R(a_, b) = 0.0; R(b, a_) = 0.0; }
            if (lane == 0) R(b, b) = sqrt(reg);
            sync();
            return 1;
        }
        proj_column(b);
        for (int a_ = lane; a_ <= b; a_ += TEAM) {
            double s = 0.0;
            for (int p = 0; p < nFR; p++) s += Q[p * ld + a_] * t2[FR[p]];
            w[a_] = s;
        }
        sync();
        // r = R'^{-1} w[0..b): forward substitution, column oriented
        for (int k = 0; k < b; k++) {
            double rk = w[k] / R(k, k);
            sync();
            if (lane == 0) R(k, b) = rk;
            for (int i = k + 1 + lane; i < b; i += TEAM) w[i] -= R(k, i) * rk;
            sync();
        }
        double rho2 = w[b];
        for (int k = 0; k < b; k++) rho2 -= R(k, b) * R(k, b);
        int ok = check_curvature ? (rho2 > QP_EPS_FLIP) : (rho2 > QP_ZERO);
        sync();  // every lane has read w[b] / R(.,b) before a caller may overwrite them (team-uniform decision)
        if (!ok) return 0;
        if (lane == 0) R(b, b) = sqrt(rho2);
        for (int a_ = lane; a_ < b; a_ += TEAM) R(b, a_) = 0.0;
        sync();
        return 1;
    }

    // ---------------------------------------------------------------- working-set updates
    __device__ QP_FN void constraint_w(int c, double& wz2, double& a2) {
        int nZ = nFR - nAC;
        for (int p = lane; p < nFR; p += TEAM) a[p] = A_entry(c, FR[p]);
        sync();
        for (int j = lane; j < nFR; j += TEAM) {
            double s = 0.0;
            for (int p = 0; p < nFR; p++) s += Q[p * ld + j] * a[p];
            w[j] = s;
        }
        sync();
        double s2 = 0.0, z2 = 0.0;
        for (int p = 0; p < nFR; p++) s2 += a[p] * a[p];
        for (int j = 0; j < nZ; j++) z2 += w[j] * w[j];
        wz2 = z2; a2 = s2;
        sync();
    }
    // chain of rotations compressing w[0..cnt) into w[cnt-1]; writes (c,s) to t2,t3; returns r
    __device__ QP_FN double rotation_chain(int cnt) {
        double r = w[0];
        if (lane == 0) {
            double a0 = w[0];
            for (int j = 0; j + 1 < cnt; j++) {
                double c, s;
                givens(a0, w[j + 1], c, s, a0);
                t2[j] = c; t3[j] = s;
            }
            t2[cnt - 1] = a0;
        }
        sync();
        r = t2[cnt - 1];
        return r;
    }
    __device__ QP_FN void add_constraint(int c, int status) {
        int nZ = nFR - nAC;
        double r = rotation_chain(nZ);
        for (int p = lane; p < nFR; p += TEAM) {
            double* q = Q + p * ld;
            double qa = q[0];
            for (int j = 0; j + 1 < nZ; j++) {
                double cs = t2[j], sn = t3[j], qb = q[j + 1];
                q[j] = cs * qa - sn * qb;
                qa = sn * qa + cs * qb;
            }
            if (nZ > 0) q[nZ - 1] = qa;
        }
        for (int j = lane; j < nFR; j += TEAM) T(nAC, j) = (j > nZ - 1) ? w[j] : ((j == nZ - 1) ? r : 0.0);
        if (lane == 0) { AC[nAC] = (short)c; posAC[c] = (short)nAC; sC[c] = (short)status; }
        nAC++;
        sync();
    }
    __device__ QP_FN void remove_constraint(int c) {
        int k = posAC[c];
        for (int i = k + 1; i < nAC; i++) {
            int cL = nFR - 1 - i;
            double cs, sn, r;
            givens(T(i, cL), T(i, cL + 1), cs, sn, r);
            sync();
            for (int ii = i + lane; ii < nAC; ii += TEAM) {
                double ta = T(ii, cL), tb = T(ii, cL + 1);
                T(ii, cL) = (ii == i) ? 0.0 : cs * ta - sn * tb;
                T(ii, cL + 1) = sn * ta + cs * tb;
            }
            for (int p = lane; p < nFR; p += TEAM) {
                double qa = Q[p * ld + cL], qb = Q[p * ld + cL + 1];
                Q[p * ld + cL] = cs * qa - sn * qb;
                Q[p * ld + cL + 1] = sn * qa + cs * qb;
            }
            sync();
        }
        // shift rows k+1.. up by one (row i -> i-1): sequential over rows, lanes over columns
        for (int i = k + 1; i < nAC; i++) {
            for (int j = lane; j < nFR; j += TEAM) T(i - 1, j) = T(i, j);
            sync();
        }
        if (lane == 0) {
            for (int i = k + 1; i < nAC; i++) { AC[i - 1] = AC[i]; posAC[AC[i - 1]] = (short)(i - 1); }
            sC[c] = 0; posAC[c] = -1;
        }
        nAC--;
        sync();
    }
    __device__ QP_FN double bound_w(int v) {
        int nZ = nFR - nAC, p = posFR[v];
        for (int j = lane; j < nFR; j += TEAM) w[j] = Q[p * ld + j];
        sync();
        double z2 = 0.0;
        for (int j = 0; j < nZ; j++) z2 += w[j] * w[j];
        sync();
        return z2;
    }
    __device__ QP_FN void add_bound(int v, int status) {
        int nZ = nFR - nAC, p = posFR[v];
        rotation_chain(nFR);
        for (int pp = lane; pp < nFR; pp += TEAM) {
            double* q = Q + pp * ld;
            double qa = q[0];
            for (int j = 0; j + 1 < nFR; j++) {
                double cs = t2[j], sn = t3[j], qb = q[j + 1];
                q[j] = cs * qa - sn * qb;
                qa = sn * qa + cs * qb;
            }
            q[nFR - 1] = qa;
        }
        // T rows: row i is touched by rotations j >= nFR-2-i (and j >= nZ-1)
        for (int i = lane; i < nAC; i += TEAM) {
            int j0 = nFR - 2 - i; if (j0 < nZ - 1) j0 = nZ - 1; if (j0 < 0) j0 = 0;
            double ta = T(i, j0);
            for (int j = j0; j + 1 < nFR; j++) {
                double cs = t2[j], sn = t3[j], tb = T(i, j + 1);
                T(i, j) = cs * ta - sn * tb;
                ta = sn * ta + cs * tb;
            }
            T(i, nFR - 1) = ta;
        }
        sync();
        int last = nFR - 1;
        if (p != last) {
            for (int j = lane; j < nFR - 1; j += TEAM) Q[p * ld + j] = Q[last * ld + j];
            if (lane == 0) { short vl = FR[last]; FR[p] = vl; posFR[vl] = (short)p; }
        }
        if (lane == 0) { posFR[v] = -1; sB[v] = (short)status; }
        nFR--;
        sync();
    }
    __device__ QP_FN void remove_bound(int v) {
        for (int j = lane; j < nFR; j += TEAM) { Q[nFR * ld + j] = 0.0; Q[j * ld + nFR] = 0.0; }
        for (int i = lane; i < nAC; i += TEAM) {
            int r = AC[i];
            double s = 0.0;
            int e1 = Ap[v + 1];
            for (int e = Ap[v]; e < e1; e++)
                if (Ai[e] == r) s += Av[e];
            T(i, nFR) = s;
        }
        if (lane == 0) { Q[nFR * ld + nFR] = 1.0; FR[nFR] = (short)v; posFR[v] = (short)nFR; sB[v] = 0; }
        nFR++;
        sync();
        for (int i = 0; i < nAC; i++) {
            int cL = nFR - 2 - i;
            double cs, sn, r;
            givens(T(i, cL), T(i, cL + 1), cs, sn, r);
            sync();
            for (int ii = i + lane; ii < nAC; ii += TEAM) {
                double ta = T(ii, cL), tb = T(ii, cL + 1);
                T(ii, cL) = (ii == i) ? 0.0 : cs * ta - sn * tb;
                T(ii, cL + 1) = sn * ta + cs * tb;
            }
            for (int p = lane; p < nFR; p += TEAM) {
                double qa = Q[p * ld + cL], qb = Q[p * ld + cL + 1];
                Q[p * ld + cL] = cs * qa - sn * qb;
                Q[p * ld + cL + 1] = sn * qa + cs * qb;
            }
            sync();
        }
    }

    // ---------------------------------------------------------------- T solves
    // T v = b : v indexed by Q column; row i has its diagonal at column nFR-1-i
    __device__ QP_FN void solve_T(double* b, double* v) {  // b (by AC position) is destroyed
        for (int i = 0; i < nAC; i++) {
            int d = nFR - 1 - i;
            double vi = b[i] / T(i, d);
            sync();
            if (lane == 0) v[d] = vi;
            for (int k = i + 1 + lane; k < nAC; k += TEAM) b[k] -= T(k, d) * vi;
            sync();
        }
    }
    // T' u = r : r indexed by Q column (destroyed), u by AC position
    __device__ QP_FN void solve_Tt(double* r, double* u) {
        for (int i = nAC - 1; i >= 0; i--) {
            int d = nFR - 1 - i;
            double ui = r[d] / T(i, d);
            sync();
            if (lane == 0) u[i] = ui;
            // r[d'] -= T(i, d') * u_i for the remaining unknowns k < i, i.e. columns d' = nFR-1-k > d
            for (int k = lane; k < i; k += TEAM) { int dk = nFR - 1 - k; r[dk] -= T(i, dk) * ui; }
            sync();
        }
    }

    // ---------------------------------------------------------------- step direction
    // dxFX: full-length vector holding the bound shift of every fixed variable; dbAC by AC position.
    // dgv(i) is evaluated on the fly as gN[i]-g[i] when use_dg, else taken from `dgvec`.
    __device__ QP_FN void step_direction(const double* dgvec, const double* dxFX, double* dbAC) {
        int nZ = nFR - nAC;
        for (int i = lane; i < nV; i += TEAM) dx[i] = (sB[i] != 0) ? dxFX[i] : 0.0;
        sync();
        if (nAC > 0) {
            mulA(dx, t2);
            for (int i = lane; i < nAC; i += TEAM) t3[i] = dbAC[i] - t2[AC[i]];
            sync();
            solve_T(t3, yv);
            for (int p = lane; p < nFR; p += TEAM) {
                double s = 0.0;
                for (int j = nZ; j < nFR; j++) s += Q[p * ld + j] * yv[j];
                dx[FR[p]] = s;
            }
            sync();
        }
        if (nZ > 0) {
            mulH(dx, t1);
            for (int j = lane; j < nZ; j += TEAM) {
                double s = 0.0;
                for (int p = 0; p < nFR; p++) { int v = FR[p]; s += Q[p * ld + j] * (t1[v] + dgvec[v]); }
                zv[j] = -s;
            }
            sync();
            // R' u = rhs (forward), R z = u (backward); column oriented
            for (int k = 0; k < nZ; k++) {
                double uk = zv[k] / R(k, k);
                sync();
                if (lane == 0) zv[k] = uk;
                for (int i = k + 1 + lane; i < nZ; i += TEAM) zv[i] -= R(k, i) * uk;
                sync();
            }
            for (int k = nZ - 1; k >= 0; k--) {
                double zk = zv[k] / R(k, k);
                sync();
                if (lane == 0) zv[k] = zk;
                for (int i = lane; i < k; i += TEAM) zv[i] -= R(i, k) * zk;
                sync();
            }
            for (int p = lane; p < nFR; p += TEAM) {
                double s = 0.0;
                for (int j = 0; j < nZ; j++) s += Q[p * ld + j] * zv[j];
                dx[FR[p]] += s;
            }
            sync();
        }
        mulH(dx, t1);
        for (int i = lane; i < nV; i += TEAM) t1[i] += dgvec[i];
        for (int i = lane; i < nT; i += TEAM) dy[i] = 0.0;
        sync();
        if (nAC > 0) {
            for (int j = nZ + lane; j < nFR; j += TEAM) {
                double s = 0.0;
                for (int p = 0; p < nFR; p++) s += Q[p * ld + j] * t1[FR[p]];
                yv[j] = s;
            }
            sync();
            solve_Tt(yv, t3);
            for (int i = lane; i < nAC; i += TEAM) dy[nV + AC[i]] = t3[i];
            sync();
            mulAT(dy + nV, t2);
            for (int i = lane; i < nV; i += TEAM) if (sB[i] != 0) dy[i] = t1[i] - t2[i];
        } else {
            for (int i = lane; i < nV; i += TEAM) if (sB[i] != 0) dy[i] = t1[i];
        }
        sync();
    }

    // ---------------------------------------------------------------- drift correction / ramping
    __device__ QP_FN void stationarity_gradient() {  // g = A'y_c + y_b - (H+regI) x
        mulAT(y + nV, t2);
        mulH(x, t1);
        for (int i = lane; i < nV; i += TEAM) g[i] = t2[i] + y[i] - t1[i];
        sync();
    }
    __device__ QP_FN void drift_correction() {
        mulA(x, Ax);
        for (int i = lane; i < nV; i += TEAM) {
            int s = sB[i]; double xi = x[i];
            if (s < 0) { lb[i] = xi; if (ub[i] < xi) ub[i] = xi; if (y[i] < 0) y[i] = 0.0; }
            else if (s > 0) { ub[i] = xi; if (lb[i] > xi) lb[i] = xi; if (y[i] > 0) y[i] = 0.0; }
            else { if (lb[i] > xi) lb[i] = xi; if (ub[i] < xi) ub[i] = xi; y[i] = 0.0; }
        }
        for (int i = lane; i < nC; i += TEAM) {
            int s = sC[i]; double ax = Ax[i];
            if (s < 0) { lbA[i] = ax; if (ubA[i] < ax) ubA[i] = ax; if (y[nV + i] < 0) y[nV + i] = 0.0; }
            else if (s > 0) { ubA[i] = ax; if (lbA[i] > ax) lbA[i] = ax; if (y[nV + i] > 0) y[nV + i] = 0.0; }
            else { if (lbA[i] > ax) lbA[i] = ax; if (ubA[i] < ax) ubA[i] = ax; y[nV + i] = 0.0; }
        }
        sync();
        stationarity_gradient();
    }
    __device__ QP_FN void ramping() {
        int nRamp = nV + nC + nC + nV;
        const double r0 = 0.5, r1 = 1.0;
        mulA(x, Ax);
        for (int i = lane; i < nV; i += TEAM) {
            double tP = (double)((i + ramp_offset) % nRamp) / (double)(nRamp - 1);
            double rP = (1.0 - tP) * r0 + tP * r1;
            double tD = (double)((nV + nC + i + ramp_offset) % nRamp) / (double)(nRamp - 1);
            double rD = (1.0 - tD) * r0 + tD * r1;
            double xi = x[i];
            double sca = fabs(xi) > 1.0 ? fabs(xi) : 1.0;
            int s = sB[i];
            if (s >= 0) lb[i] = xi - sca * rP;
            if (s <= 0) ub[i] = xi + sca * rP;
            if (s < 0) { lb[i] = xi; y[i] = rD; }
            if (s > 0) { ub[i] = xi; y[i] = -rD; }
            if (s == 0) y[i] = 0.0;
        }
        for (int i = lane; i < nC; i += TEAM) {
            double tP = (double)((nV + i + ramp_offset) % nRamp) / (double)(nRamp - 1);
            double rP = (1.0 - tP) * r0 + tP * r1;
            double tD = (double)((nV + nC + nV + i + ramp_offset) % nRamp) / (double)(nRamp - 1);
            double rD = (1.0 - tD) * r0 + tD * r1;
            double ax = Ax[i];
            double sca = fabs(ax) > 1.0 ? fabs(ax) : 1.0;
            int s = sC[i];
            if (s >= 0) lbA[i] = ax - sca * rP;
            if (s <= 0) ubA[i] = ax + sca * rP;
            if (s < 0) { lbA[i] = ax; y[nV + i] = rD; }
            if (s > 0) { ubA[i] = ax; y[nV + i] = -rD; }
            if (s == 0) y[nV + i] = 0.0;
        }
        sync();
        stationarity_gradient();
        ramp_offset++;
    }

    // ---------------------------------------------------------------- exchange (ensure LI)
    // element to add: constraint c (v<0) or bound v (c<0); w holds its Q-coordinates. 0 ok, 1 infeasible
    __device__ QP_FN int ensure_li(int c, int v, int status) {
        int nZ = nFR - nAC;
        double* xiC = zv;
        double* xiB = dx;
        for (int j = nZ + lane; j < nFR; j += TEAM) yv[j] = w[j];
        sync();
        solve_Tt(yv, xiC);
        for (int i = lane; i < nC; i += TEAM) { int p = posAC[i]; t3[i] = (p >= 0) ? xiC[p] : 0.0; }
        sync();
        mulAT(t3, t2);
        for (int i = lane; i < nV; i += TEAM) {
            if (sB[i] == 0) { xiB[i] = 0.0; continue; }
            double ai = (c >= 0) ? A_entry(c, i) : 0.0;
            xiB[i] = ai - t2[i];
        }
        sync();
        double sgn = (status < 0) ? 1.0 : -1.0;
        double best = QP_INFTY; int bpos = 0x7fffffff;
        for (int i = lane; i < nAC; i += TEAM) {
            int ci = AC[i]; double xi = sgn * xiC[i], yy = y[nV + ci]; double t = QP_INFTY;
            if (sC[ci] < 0) { if (xi > QP_ZERO && yy >= 0.0) t = yy / xi; }
            else { if (xi < -QP_ZERO && yy <= 0.0) t = yy / xi; }
            if (t < QP_INFTY && key_less(t, i, best, bpos)) { best = t; bpos = i; }
        }
        for (int i = lane; i < nV; i += TEAM) {
            if (sB[i] == 0) continue;
            double xi = sgn * xiB[i], yy = y[i]; double t = QP_INFTY;
            if (sB[i] < 0) { if (xi > QP_ZERO && yy >= 0.0) t = yy / xi; }
            else { if (xi < -QP_ZERO && yy <= 0.0) t = yy / xi; }
            if (t < QP_INFTY && key_less(t, nC + i, best, bpos)) { best = t; bpos = nC + i; }
        }
        MinKey mk = team_min(best, bpos);
        if (mk.pos == 0x7fffffff) return 1;
        double ymin = mk.t;
        int kind = mk.pos >= nC ? 1 : 0;
        int idx = kind ? mk.pos - nC : AC[mk.pos];
        for (int i = lane; i < nAC; i += TEAM) y[nV + AC[i]] -= ymin * sgn * xiC[i];
        for (int i = lane; i < nV; i += TEAM) if (sB[i] != 0) y[i] -= ymin * sgn * xiB[i];
        sync();
        if (lane == 0) {
            if (c >= 0) y[nV + c] = sgn * ymin; else y[v] = sgn * ymin;
            if (kind == 0) y[nV + idx] = 0.0; else y[idx] = 0.0;
        }
        sync();
        if (kind == 0) remove_constraint(idx); else remove_bound(idx);
        return 0;
    }

    // ---------------------------------------------------------------- homotopy
    __device__ QP_FN int homotopy(int max_iter) {
        iters = 0;
        for (int it = 0;; it++) {
            // data shift: dg in `a`.. we keep dg in zv? -> use dedicated: t-vectors are busy, so dg lives in `a` (nT)
            // w <- bound shift of fixed variables, yv <- constraint shift by AC position, a <- dg
            for (int i = lane; i < nV; i += TEAM) {
                int s = sB[i];
                w[i] = s < 0 ? (lbN[i] - lb[i]) : (s > 0 ? (ubN[i] - ub[i]) : 0.0);
                a[i] = gN[i] - g[i];
            }
            sync();
            // dbAC goes to the tail of `w` (entries nV..nV+nAC), nAC <= nC
            for (int i = lane; i < nAC; i += TEAM) { int ci = AC[i]; w[nV + i] = sC[ci] < 0 ? (lbAN[ci] - lbA[ci]) : (ubAN[ci] - ubA[ci]); }
            sync();
            step_direction(a, w, w + nV);
            mulA(dx, dAx);

            // ---- ratio tests: position = scan order of the sequential rule
            double best = 2.0; int bpos = 0x7fffffff;
#define CONSIDER(num, den, pos)                                                                   \
    {                                                                                             \
        double n_ = (num), d_ = (den);                                                            \
        if (d_ >= QP_EPS_DEN && n_ >= QP_EPS_NUM && n_ < d_) {                                    \
            double t_ = n_ / d_;                                                                  \
            if (t_ < 1.0 && key_less(t_, (pos), best, bpos)) { best = t_; bpos = (pos); }         \
        }                                                                                         \
    }
            for (int i = lane; i < nAC; i += TEAM) {
                int ci = AC[i];
                if (sC[ci] < 0) CONSIDER(y[nV + ci], -dy[nV + ci], i) else CONSIDER(-y[nV + ci], dy[nV + ci], i)
            }
            for (int i = lane; i < nV; i += TEAM) {
                int s = sB[i];
                if (s == 0) continue;
                if (s < 0) CONSIDER(y[i], -dy[i], nC + i) else CONSIDER(-y[i], dy[i], nC + i)
            }
            for (int i = lane; i < nC; i += TEAM) {
                if (sC[i] != 0) continue;
                double num = Ax[i] - lbA[i]; if (num < 0) num = 0;
                CONSIDER(num, (lbAN[i] - lbA[i]) - dAx[i], nC + nV + i)
                num = ubA[i] - Ax[i]; if (num < 0) num = 0;
                CONSIDER(num, dAx[i] - (ubAN[i] - ubA[i]), nC + nV + nC + i)
            }
            for (int i = lane; i < nV; i += TEAM) {
                if (sB[i] != 0) continue;
                double num = x[i] - lb[i]; if (num < 0) num = 0;
                CONSIDER(num, (lbN[i] - lb[i]) - dx[i], 2 * nC + nV + nC + i)
                num = ub[i] - x[i]; if (num < 0) num = 0;
                CONSIDER(num, dx[i] - (ubN[i] - ub[i]), 3 * nC + 2 * nV + i)
            }
#undef CONSIDER
            MinKey mk = team_min(best, bpos);
            double tau = 1.0; int bc_idx = -1, bc_isbound = 0, bc_status = 0;
            if (mk.pos != 0x7fffffff) {
                tau = mk.t;
                int p = mk.pos;
                if (p < nC) { bc_idx = AC[p]; bc_isbound = 0; bc_status = 0; }
                else if (p < nC + nV) { bc_idx = p - nC; bc_isbound = 1; bc_status = 0; }
                else if (p < 2 * nC + nV) { bc_idx = p - nC - nV; bc_isbound = 0; bc_status = -1; }
                else if (p < 3 * nC + nV) { bc_idx = p - 2 * nC - nV; bc_isbound = 0; bc_status = 1; }
                else if (p < 3 * nC + 2 * nV) { bc_idx = p - 3 * nC - nV; bc_isbound = 1; bc_status = -1; }
                else { bc_idx = p - 3 * nC - 2 * nV; bc_isbound = 1; bc_status = 1; }
            }
            // ---- step
            if (bc_idx < 0) {
                for (int i = lane; i < nV; i += TEAM) { x[i] += dx[i]; g[i] = gN[i]; lb[i] = lbN[i]; ub[i] = ubN[i]; }
                for (int i = lane; i < nT; i += TEAM) y[i] += dy[i];
                for (int i = lane; i < nC; i += TEAM) { Ax[i] += dAx[i]; lbA[i] = lbAN[i]; ubA[i] = ubAN[i]; }
                sync();
                iters = it;
                return ST_OPTIMAL;
            }
            if (it >= max_iter) { iters = it; return ST_HOMOTOPY; }
            if (tau > 0.0) {
                for (int i = lane; i < nV; i += TEAM) {
                    x[i] += tau * dx[i]; g[i] += tau * a[i];
                    lb[i] += tau * (lbN[i] - lb[i]); ub[i] += tau * (ubN[i] - ub[i]);
                }
                for (int i = lane; i < nT; i += TEAM) y[i] += tau * dy[i];
                for (int i = lane; i < nC; i += TEAM) {
                    Ax[i] += tau * dAx[i];
                    lbA[i] += tau * (lbAN[i] - lbA[i]); ubA[i] += tau * (ubAN[i] - ubA[i]);
                }
                sync();
            }
            // ---- change the working set
            if (bc_status == 0) {
                int flipped = 0;
                if (bc_isbound) {
                    int old = sB[bc_idx];
                    sync();
                    if (lane == 0) y[bc_idx] = 0.0;
                    remove_bound(bc_idx);
                    if (!extend_R(flags & FLAG_FLIPPING)) {
                        if (!(flags & FLAG_FLIPPING)) { iters = it; return ST_UNBOUNDED; }
                        bound_w(bc_idx);
                        add_bound(bc_idx, -old);
                        if (lane == 0) { if (old < 0) ub[bc_idx] = x[bc_idx]; else lb[bc_idx] = x[bc_idx]; }
                        sync();
                        flipped = 1;
                    }
                } else {
                    int old = sC[bc_idx];
                    sync();
                    if (lane == 0) y[nV + bc_idx] = 0.0;
                    remove_constraint(bc_idx);
                    if (!extend_R(flags & FLAG_FLIPPING)) {
                        if (!(flags & FLAG_FLIPPING)) { iters = it; return ST_UNBOUNDED; }
                        double z2, a2;
                        constraint_w(bc_idx, z2, a2);
                        add_constraint(bc_idx, -old);
                        if (lane == 0) { if (old < 0) ubA[bc_idx] = Ax[bc_idx]; else lbA[bc_idx] = Ax[bc_idx]; }
                        sync();
                        flipped = 1;
                    }
                }
                if (flipped) {
                    if (recompute_R()) { iters = it; return ST_INTERNAL; }
                }
            } else {
                if (bc_isbound) {
                    double z2 = bound_w(bc_idx);
                    if (!(z2 > QP_EPS_LI * QP_EPS_LI)) {
                        if (ensure_li(-1, bc_idx, bc_status)) { iters = it; return ST_INFEASIBLE; }
                        bound_w(bc_idx);
                    }
                    add_bound(bc_idx, bc_status);
                } else {
                    double z2, a2;
                    constraint_w(bc_idx, z2, a2);
                    if (!(z2 > QP_EPS_LI * QP_EPS_LI * a2) || a2 == 0.0) {
                        if (ensure_li(bc_idx, -1, bc_status)) { iters = it; return ST_INFEASIBLE; }
                        constraint_w(bc_idx, z2, a2);
                    }
                    add_constraint(bc_idx, bc_status);
                }
                if (recompute_R()) { iters = it; return ST_INTERNAL; }
            }
            if (tau <= QP_EPS && (flags & FLAG_RAMPING)) ramping();
            else if (flags & FLAG_DRIFT) drift_correction();
        }
    }

    // ---------------------------------------------------------------- cold start / refactorise
    __device__ QP_FN void cold_start_state() {
        nFR = 0; nAC = 0; ramp_offset = 0;
        for (int i = lane; i < nV; i += TEAM) {
            x[i] = 0.0; y[i] = 0.0; sB[i] = -1; posFR[i] = -1; g[i] = 0.0; lb[i] = 0.0; ub[i] = QP_BOUND_RELAX;
        }
        for (int i = lane; i < nC; i += TEAM) {
            y[nV + i] = 0.0; sC[i] = 0; posAC[i] = -1; Ax[i] = 0.0; lbA[i] = -QP_BOUND_RELAX; ubA[i] = QP_BOUND_RELAX;
        }
        sync();
    }
    // rebuild TQ and R for the kept working set with the new matrix values; 0 ok
    __device__ QP_FN int refactorise() {
        int nAC_old = nAC;
        // remember (constraint, status) by AC position in t1/t3 tails is unsafe (used by callees): use yv/zv? also used.
        // -> keep them in dAx (nC) and dy[nV..] (nC): neither is touched by constraint_w/add_constraint/recompute_R.
        for (int i = lane; i < nAC_old; i += TEAM) { int ci = AC[i]; dAx[i] = (double)ci; dy[nV + i] = (double)sC[ci]; }
        sync();
        for (int k = lane; k < nFR * nFR; k += TEAM) { int i = k / nFR, j = k % nFR; Q[i * ld + j] = (i == j) ? 1.0 : 0.0; }
        for (int i = lane; i < nAC_old; i += TEAM) { int ci = (int)dAx[i]; sC[ci] = 0; posAC[ci] = -1; }
        nAC = 0;
        sync();
        for (int i = 0; i < nAC_old; i++) {
            int ci = (int)dAx[i]; int st = (int)dy[nV + i];
            double z2, a2;
            constraint_w(ci, z2, a2);
            if (!(z2 > QP_EPS_LI * QP_EPS_LI * a2) || a2 == 0.0) { sync(); if (lane == 0) y[nV + ci] = 0.0; sync(); continue; }
            add_constraint(ci, st);
        }
        return recompute_R();
    }
};

// -------------------------------------------------------------------------------------------
// kernel
// -------------------------------------------------------------------------------------------
template <int TEAM, int CTA_THREADS>
__global__ void __launch_bounds__(CTA_THREADS) qp_solve_kernel(const QPKernelArgs A) {
    extern __shared__ __align__(16) double smem[];
    constexpr int TEAMS = CTA_THREADS / TEAM;
    const int team_id = threadIdx.x / TEAM;
    const int lane = threadIdx.x % TEAM;
    const int b = blockIdx.x * TEAMS + team_id;
    if (b >= A.batch) return;
    if (A.mask && !A.mask[b]) return;

    QPSolver<TEAM> S;
    S.tm.lane = lane; S.tm.team_id = team_id; S.lane = lane;
    S.nV = A.nV; S.nC = A.nC; S.nT = A.nV + A.nC; S.ld = A.ld;
    S.is_lp = A.is_lp; S.has_H = A.has_H && !A.is_lp; S.flags = A.flags;
    S.reg = A.is_lp ? QP_EPS_REG : 0.0;
    S.Ap = A.Ap; S.Ai = A.Ai; S.Arp = A.Arp; S.Aci = A.Aci; S.Aperm = A.Aperm; S.Hp = A.Hp; S.Hi = A.Hi;
    const int nV = A.nV, nC = A.nC, nT = nV + nC;
    double* slice = smem + (size_t)team_id * A.slice_doubles;
    S.carve(slice, A);
    S.red = smem + (size_t)TEAMS * A.slice_doubles + team_id * 2 * (TEAM / 32 + 1);
    S.gN = A.gN + (size_t)b * nV; S.lbN = A.lbN + (size_t)b * nV; S.ubN = A.ubN + (size_t)b * nV;
    S.lbAN = A.lbAN + (size_t)b * nC; S.ubAN = A.ubAN + (size_t)b * nC;

    int mode = A.inst_mode ? A.inst_mode[b] : A.mode;
    int* hdr = A.state_hdr ? A.state_hdr + (size_t)b * 4 : nullptr;
    if (mode != MODE_COLD && (!hdr || !hdr[3])) mode = MODE_COLD;

    if (mode != MODE_COLD) {
        const double* st = A.state + (size_t)b * A.slice_doubles;
        for (int i = lane; i < A.slice_doubles; i += TEAM) slice[i] = st[i];
        S.nFR = hdr[0]; S.nAC = hdr[1]; S.ramp_offset = hdr[2];
    }
    if (mode != MODE_HOT_FIXED) {
        const double* av = A.Aval + (size_t)b * A.zA;
        for (int i = lane; i < A.zA; i += TEAM) S.Av[i] = av[i];
        if (S.has_H) {
            const double* hv = A.Hval + (size_t)b * A.zH;
            for (int i = lane; i < A.zH; i += TEAM) S.Hv[i] = hv[i];
        }
    }
    S.sync();

    int status;
    int total_iters = 0;
    if (mode == MODE_HOT_VARIED) {
        if (S.refactorise()) mode = MODE_COLD;  // projected Hessian of the kept set not PD: cold start
        else { S.drift_correction(); }
    }
    if (mode == MODE_COLD) S.cold_start_state();
    status = S.homotopy(A.max_iter);
    total_iters += S.iters;
    if (status != ST_OPTIMAL && mode != MODE_COLD) {
        // one-retry recovery of handle_error (src/qpOASESInterface.cpp:746-754): plain re-init
        S.cold_start_state();
        status = S.homotopy(A.max_iter);
        total_iters += S.iters;
    }

    // ---------------- epilogue: results, working set, fused KKT residuals (test_optimality)
    double* xo = A.x + (size_t)b * nV;
    double* yo = A.y + (size_t)b * nT;
    for (int i = lane; i < nV; i += TEAM) xo[i] = S.x[i];
    for (int i = lane; i < nT; i += TEAM) yo[i] = S.y[i];
    if (A.wsB) for (int i = lane; i < nV; i += TEAM) A.wsB[(size_t)b * nV + i] = (signed char)S.sB[i];
    if (A.wsC) for (int i = lane; i < nC; i += TEAM) A.wsC[(size_t)b * nC + i] = (signed char)S.sC[i];

    // objective 1/2 x'Hx + g'x with the unregularised H; Hx kept in t1 for the KKT residuals
    {
        double reg_save = S.reg; S.reg = 0.0;
        S.mulH(S.x, S.t1);
        S.reg = reg_save;
        S.mulA(S.x, S.Ax);
        S.mulAT(S.y + nV, S.t2);
    }
    if (lane == 0) {
        const double SQRT_M_EPS = 1.0e-8;
        double obj = 0.0;
        for (int i = 0; i < nV; i++) obj += 0.5 * S.x[i] * S.t1[i];
        for (int i = 0; i < nV; i++) obj += S.gN[i] * S.x[i];
        A.obj[b] = obj; A.status[b] = status; A.iters[b] = total_iters;
        // qpOASESInterface::get_working_set + test_optimality (src/qpOASESInterface.cpp:835-895, 498-684),
        // evaluated against the target data the caller supplied.
        double primal = 0.0, dual = 0.0, compl_ = 0.0, stat = 0.0;
        int* WB = A.WB ? A.WB + (size_t)b * nV : nullptr;
        int* WC = A.WC ? A.WC + (size_t)b * nC : nullptr;
        for (int i = 0; i < nV; i++) {
            double xi = S.x[i], l = S.lbN[i], u = S.ubN[i];
            primal += fmax(0.0, l - xi);
            primal += -fmin(0.0, u - xi);
        }
        for (int i = 0; i < nC; i++) {
            double ax = S.Ax[i];
            primal += fmax(0.0, S.lbAN[i] - ax);
            primal += -fmin(0.0, S.ubAN[i] - ax);
        }
        for (int i = 0; i < nV; i++) {
            int s = S.sB[i], W; double xi = S.x[i], yi = S.y[i];
            if (s > 0) W = (fabs(xi - S.lbN[i]) < SQRT_M_EPS) ? -99 : 1;
            else if (s < 0) W = (fabs(xi - S.ubN[i]) < SQRT_M_EPS) ? -99 : -1;
            else W = 0;
            if (WB) WB[i] = W;
            if (W == 0) dual += fabs(yi); else if (W == -1) dual += -fmin(0.0, yi); else if (W == 1) dual += fmax(0.0, yi);
        }
        for (int i = 0; i < nC; i++) {
            int s = S.sC[i], W; double ax = S.Ax[i], yi = S.y[nV + i];
            if (s > 0) W = (ax - S.lbAN[i] < SQRT_M_EPS) ? -99 : 1;      // :874 (comparison inside fabs)
            else if (s < 0) W = (ax - S.ubAN[i] < SQRT_M_EPS) ? -99 : -1; // :880
            else W = 0;
            if (WC) WC[i] = W;
            if (W == 0) dual += fabs(yi); else if (W == -1) dual += -fmin(0.0, yi); else if (W == 1) dual += fmax(0.0, yi);
        }
        for (int i = 0; i < nV; i++) {
            double gap = S.t2[i];
            gap += S.y[i]; gap -= S.gN[i]; gap -= S.t1[i];
            stat += fabs(gap);
        }
        for (int i = 0; i < nV; i++) {
            int s = S.sB[i]; double xi = S.x[i], yi = S.y[i];
            int W = (s > 0) ? ((fabs(xi - S.lbN[i]) < SQRT_M_EPS) ? -99 : 1) : (s < 0 ? ((fabs(xi - S.ubN[i]) < SQRT_M_EPS) ? -99 : -1) : 0);
            if (W == 0) compl_ += fabs(yi); else if (W == -1) compl_ += fabs(yi * (xi - S.lbN[i])); else if (W == 1) compl_ += fabs(yi * (S.ubN[i] - xi));
        }
        for (int i = 0; i < nC; i++) {
            int s = S.sC[i]; double ax = S.Ax[i], yi = S.y[nV + i];
            int W = (s > 0) ? ((ax - S.lbAN[i] < SQRT_M_EPS) ? -99 : 1) : (s < 0 ? ((ax - S.ubAN[i] < SQRT_M_EPS) ? -99 : -1) : 0);
            if (W == 0) compl_ += fabs(yi); else if (W == -1) compl_ += fabs(yi * (ax - S.lbAN[i])); else if (W == 1) compl_ += fabs(yi * (S.ubAN[i] - ax));
        }
        if (A.kkt) {
            double* k = A.kkt + (size_t)b * 5;
            k[0] = primal; k[1] = dual; k[2] = stat; k[3] = compl_; k[4] = compl_ + stat + dual + primal;
        }
    }
    S.sync();
    if ((A.flags & FLAG_KEEP_STATE) && A.state) {
        double* st = A.state + (size_t)b * A.slice_doubles;
        for (int i = lane; i < A.slice_doubles; i += TEAM) st[i] = slice[i];
        if (lane == 0) { hdr[0] = S.nFR; hdr[1] = S.nAC; hdr[2] = S.ramp_offset; hdr[3] = (status == ST_OPTIMAL) ? 1 : 0; }
    }
}

}  // namespace sqpb200
