// l0_kernels.cuh -- sm_100a kernels for the L0 rows of SURVEY.md 8a (A4-A7), the L2 data
// construction (B2, B3) and the stand-alone KKT residual kernel (C7, C8).
//
// All of these are HBM-bound integer / FP64 streaming work: no tensor cores.  Layout is
// instance-major ([batch][len]); the sparsity pattern is shared by the batch and stays in L1/L2.
// FP64 sums that the reference evaluates with separate multiply and add
// (src/SpHbMat.cpp:734, 693) use __dmul_rn/__dadd_rn so that no FMA contraction changes a bit:
// results are bit-identical with the reference's SpHbMat::times / transposed_times.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqpb200 {

// ------------------------------------------------------------------------------------------
// A4/A5: segmented triplet -> CSC.  One CTA per matrix.  Keys (col,row,counter) are packed into
// 64 bits, so a plain (unstable) bitonic network yields the stable order of the reference's
// comparator (include/sqphot/SpHbMat.hpp:370-380) with the entry counter as tie-break.
// ------------------------------------------------------------------------------------------
#define ASM_SMEM_KEYS 4096  // keys sorted in shared memory up to this (padded) size, else in global scratch

__device__ __forceinline__ uint64_t pack_key(int row1, int col1, int cnt) {
    return ((uint64_t)(uint32_t)(col1 - 1) << 42) | ((uint64_t)(uint32_t)(row1 - 1) << 21) | (uint64_t)(uint32_t)cnt;
}

__global__ void __launch_bounds__(256) csc_assemble_kernel(int nmat, const int* __restrict__ seg,
                                                            const int* __restrict__ ncol_arr,
                                                            const int* __restrict__ cp_off,
                                                            const int* __restrict__ row1,
                                                            const int* __restrict__ col1,
                                                            const int* __restrict__ pad_off,
                                                            uint64_t* __restrict__ scratch,
                                                            int* __restrict__ colptr, int* __restrict__ rowidx,
                                                            int* __restrict__ order) {
    __shared__ uint64_t skeys[ASM_SMEM_KEYS];
    const int m = blockIdx.x;
    if (m >= nmat) return;
    const int z0 = seg[m], z = seg[m + 1] - z0, ncol = ncol_arr[m];
    const int P = pad_off[m + 1] - pad_off[m];  // power of two >= z (>= 1)
    uint64_t* keys = (P <= ASM_SMEM_KEYS) ? skeys : (scratch + pad_off[m]);
    int* cp = colptr + cp_off[m];
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        keys[i] = (i < z) ? pack_key(row1[z0 + i], col1[z0 + i], i) : ~0ull;
    __syncthreads();
    // bitonic sort, ascending
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                int l = i ^ j;
                if (l > i) {
                    uint64_t a = keys[i], b = keys[l];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    // emit rowidx, order (inverse permutation) and colptr (exclusive counts) -- src/SpHbMat.cpp:255-264
    for (int i = threadIdx.x; i < z; i += blockDim.x) {
        uint64_t kk = keys[i];
        int c = (int)(kk >> 42), r = (int)((kk >> 21) & 0x1fffff), cnt = (int)(kk & 0x1fffff);
        rowidx[z0 + i] = r;
        order[z0 + cnt] = i;
        int cprev = (i == 0) ? -1 : (int)(keys[i - 1] >> 42);
        for (int j = cprev + 1; j <= c; j++) cp[j] = i;
    }
    int clast = (z == 0) ? -1 : (int)(keys[z - 1] >> 42);
    for (int j = clast + 1 + threadIdx.x; j <= ncol; j += blockDim.x) cp[j] = z;
}

// ------------------------------------------------------------------------------------------
// A6: value refresh.  Gather form of SpHbMat::setMatVal (src/SpHbMat.cpp:368-393):
// out[b][k] = in[b][src[k]] for every CSC slot k whose source triplet is src[k] >= 0 (identity
// slots keep their +-1).  Consecutive threads write consecutive addresses.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_values_kernel(long long total, int z_in, int z_out,
                                                              const int* __restrict__ src,
                                                              const double* __restrict__ in, int broadcast,
                                                              double* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        long long b = t / z_out;
        int k = (int)(t - b * z_out);
        int s = src[k];
        if (s >= 0) out[t] = broadcast ? in[s] : in[b * z_in + s];
    }
}

// fill out[b][k] = vals[k] (initial CSC values: identity entries, zeros elsewhere; or broadcast rows)
__global__ void __launch_bounds__(256) broadcast_rows_kernel(long long total, int len, int offset, int count,
                                                              const double* __restrict__ vals, double* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        long long b = t / count;
        int k = (int)(t - b * count);
        out[b * len + offset + k] = vals[k];
    }
}

// ------------------------------------------------------------------------------------------
// A7: batched SpMV / SpMTV on a shared pattern.  One thread per (instance, output entry):
// out[b][o] = sum_{k in [ptr[o], ptr[o+1])} val[b][perm ? perm[k] : k] * x[b][idx[k]],
// accumulated in k order = the reference's storage order for that output entry.
//   A x : CSR view (ptr=Arp, idx=Aci, perm=Aperm)     A'y : CSC (ptr=Ap, idx=Ai)
//   H x : CSC of the symmetric H (row c == column c)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spmv_kernel(long long total, int nout, int nin, int nnz,
                                                    const int* __restrict__ ptr, const int* __restrict__ idx,
                                                    const int* __restrict__ perm, const double* __restrict__ val,
                                                    const double* __restrict__ x, double* __restrict__ y) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        long long b = t / nout;
        int o = (int)(t - b * nout);
        const double* vb = val + b * nnz;
        const double* xb = x + b * nin;
        double s = 0.0;
        int k1 = ptr[o + 1];
        for (int k = ptr[o]; k < k1; k++) {
            int e = perm ? perm[k] : k;
            s = __dadd_rn(s, __dmul_rn(vb[e], xb[idx[k]]));
        }
        y[t] = s;
    }
}

// ------------------------------------------------------------------------------------------
// B2/B3: batched QPhandler data construction (src/QPhandler.cpp:185-201, 358-367, 559-564,
// 287-292, 439-440, 461-462).  One thread per (instance, entry).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qphandler_bounds_kernel(int batch, int mode, int n, int m, double inf,
                                                                const double* __restrict__ delta,
                                                                const double* __restrict__ x_l, const double* __restrict__ x_u,
                                                                const double* __restrict__ x_k, const double* __restrict__ c_l,
                                                                const double* __restrict__ c_u, const double* __restrict__ c_k,
                                                                double* __restrict__ lb, double* __restrict__ ub,
                                                                double* __restrict__ lbA, double* __restrict__ ubA) {
    const int nV = n + 2 * m;
    long long total = (long long)batch * nV;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        long long b = t / nV;
        int i = (int)(t - b * nV);
        if (i < n) {
            double d = delta[b];
            double lo = x_l[b * n + i] - x_k[b * n + i], hi = x_u[b * n + i] - x_k[b * n + i];
            lb[t] = lo > -d ? lo : -d;  // std::max(x_l - x_k, -delta)
            ub[t] = hi < d ? hi : d;    // std::min(x_u - x_k,  delta)
        } else if (mode == 0) {
            ub[t] = inf;  // slack lower bounds stay at their zero initialisation
        }
        if (i < m && mode != 2) {
            lbA[b * m + i] = c_l[b * m + i] - c_k[b * m + i];
            // mode 1 = update_bounds of the non-QORE branch: ubA is never refreshed (src/QPhandler.cpp:358-360);
            // mode 3 = update_bounds refreshing both sides, as the QORE branch does (:377-382)
            if (mode == 0 || mode == 3) ubA[b * m + i] = c_u[b * m + i] - c_k[b * m + i];
        }
    }
}

__global__ void __launch_bounds__(256) qphandler_g_kernel(int batch, int n, int m, const double* __restrict__ grad,
                                                           const double* __restrict__ rho, double* __restrict__ g) {
    const int nV = n + 2 * m;
    long long total = (long long)batch * nV;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        long long b = t / nV;
        int i = (int)(t - b * nV);
        if (i < n) { if (grad) g[t] = grad[b * n + i]; }
        else if (rho) g[t] = rho[b];
    }
}

// ------------------------------------------------------------------------------------------
// C7 + C8 stand-alone: working-set translation and KKT residuals, one warp per instance.
// Lanes evaluate A x, A'y_c and H x (one lane per output entry, reference order); lane 0 then
// accumulates the four violation sums in the reference's index order, so the result is
// bit-identical with qpOASESInterface::test_optimality (src/qpOASESInterface.cpp:498-684).
// Dynamic shared memory: per warp (2*nV + nC) doubles.
// ------------------------------------------------------------------------------------------
struct KKTArgs {
    int batch, nV, nC, zA, zH, has_H;
    const int *Ap, *Ai, *Arp, *Aci, *Aperm, *Hp, *Hi;
    const double *Aval, *Hval, *g, *lb, *ub, *lbA, *ubA, *x, *y;
    const signed char *wsB, *wsC;
    int *WB, *WC;
    double* out;
};

__global__ void __launch_bounds__(128) kkt_kernel(const KKTArgs A) {
    extern __shared__ __align__(16) double ksm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.batch) return;
    const int nV = A.nV, nC = A.nC;
    double* Ax = ksm + (size_t)warp * (2 * nV + nC);
    double* ATy = Ax + nC;
    double* Hx = ATy + nV;
    const double* av = A.Aval + b * A.zA;
    const double* hv = A.has_H ? A.Hval + b * A.zH : nullptr;
    const double *x = A.x + b * nV, *y = A.y + b * (nV + nC), *g = A.g + b * nV;
    const double *lb = A.lb + b * nV, *ub = A.ub + b * nV, *lbA = A.lbA + b * nC, *ubA = A.ubA + b * nC;
    for (int r = lane; r < nC; r += 32) {
        double s = 0.0;
        for (int k = A.Arp[r]; k < A.Arp[r + 1]; k++) s = __dadd_rn(s, __dmul_rn(av[A.Aperm[k]], x[A.Aci[k]]));
        Ax[r] = s;
    }
    for (int c = lane; c < nV; c += 32) {
        double s = 0.0;
        for (int e = A.Ap[c]; e < A.Ap[c + 1]; e++) s = __dadd_rn(s, __dmul_rn(av[e], y[nV + A.Ai[e]]));
        ATy[c] = s;
        double h = 0.0;
        if (hv)
            for (int e = A.Hp[c]; e < A.Hp[c + 1]; e++) h = __dadd_rn(h, __dmul_rn(hv[e], x[A.Hi[e]]));
        Hx[c] = h;
    }
    __syncwarp();
    if (lane != 0) return;
    const double SQRT_M_EPS = 1.0e-8;
    const signed char* sB = A.wsB + b * nV;
    const signed char* sC = A.wsC + b * nC;
    int* WB = A.WB ? A.WB + b * nV : nullptr;
    int* WC = A.WC ? A.WC + b * nC : nullptr;
    double primal = 0.0, dual = 0.0, stat = 0.0, cmpl = 0.0;
    for (int i = 0; i < nV; i++) {
        primal = __dadd_rn(primal, fmax(0.0, __dsub_rn(lb[i], x[i])));
        primal = __dadd_rn(primal, -fmin(0.0, __dsub_rn(ub[i], x[i])));
    }
    for (int i = 0; i < nC; i++) {
        primal = __dadd_rn(primal, fmax(0.0, __dsub_rn(lbA[i], Ax[i])));
        primal = __dadd_rn(primal, -fmin(0.0, __dsub_rn(ubA[i], Ax[i])));
    }
#define WSB(i) ((sB[i] > 0) ? ((fabs(__dsub_rn(x[i], lb[i])) < SQRT_M_EPS) ? -99 : 1) \
                            : ((sB[i] < 0) ? ((fabs(__dsub_rn(x[i], ub[i])) < SQRT_M_EPS) ? -99 : -1) : 0))
#define WSC(i) ((sC[i] > 0) ? ((__dsub_rn(Ax[i], lbA[i]) < SQRT_M_EPS) ? -99 : 1) \
                            : ((sC[i] < 0) ? ((__dsub_rn(Ax[i], ubA[i]) < SQRT_M_EPS) ? -99 : -1) : 0))
    for (int i = 0; i < nV; i++) {
        int W = WSB(i); double yi = y[i];
        if (WB) WB[i] = W;
        if (W == 0) dual = __dadd_rn(dual, fabs(yi));
        else if (W == -1) dual = __dadd_rn(dual, -fmin(0.0, yi));
        else if (W == 1) dual = __dadd_rn(dual, fmax(0.0, yi));
    }
    for (int i = 0; i < nC; i++) {
        int W = WSC(i); double yi = y[nV + i];
        if (WC) WC[i] = W;
        if (W == 0) dual = __dadd_rn(dual, fabs(yi));
        else if (W == -1) dual = __dadd_rn(dual, -fmin(0.0, yi));
        else if (W == 1) dual = __dadd_rn(dual, fmax(0.0, yi));
    }
    for (int i = 0; i < nV; i++) {
        double gap = ATy[i];
        gap = __dadd_rn(gap, y[i]);
        gap = __dsub_rn(gap, g[i]);
        gap = __dsub_rn(gap, Hx[i]);
        stat = __dadd_rn(stat, fabs(gap));
    }
    for (int i = 0; i < nV; i++) {
        int W = WSB(i); double yi = y[i];
        if (W == 0) cmpl = __dadd_rn(cmpl, fabs(yi));
        else if (W == -1) cmpl = __dadd_rn(cmpl, fabs(__dmul_rn(yi, __dsub_rn(x[i], lb[i]))));
        else if (W == 1) cmpl = __dadd_rn(cmpl, fabs(__dmul_rn(yi, __dsub_rn(ub[i], x[i]))));
    }
    for (int i = 0; i < nC; i++) {
        int W = WSC(i); double yi = y[nV + i];
        if (W == 0) cmpl = __dadd_rn(cmpl, fabs(yi));
        else if (W == -1) cmpl = __dadd_rn(cmpl, fabs(__dmul_rn(yi, __dsub_rn(Ax[i], lbA[i]))));
        else if (W == 1) cmpl = __dadd_rn(cmpl, fabs(__dmul_rn(yi, __dsub_rn(ubA[i], Ax[i]))));
    }
#undef WSB
#undef WSC
    double* o = A.out + b * 5;
    o[0] = primal; o[1] = dual; o[2] = stat; o[3] = cmpl;
    o[4] = __dadd_rn(__dadd_rn(__dadd_rn(cmpl, stat), dual), primal);
}

}  // namespace sqpb200
