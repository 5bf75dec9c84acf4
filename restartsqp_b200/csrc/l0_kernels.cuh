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
#include "tma.cuh"

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

// One warp per matrix (every padded size <= ASM_WARP_KEYS): the bitonic network runs out of the warp's own shared-memory
// slice with warp barriers only, eight matrices per CTA.  Same keys, same outputs as csc_assemble_kernel.
#define ASM_WARP_KEYS 512
__global__ void __launch_bounds__(256) csc_assemble_warp_kernel(int nmat, int Pmax, const int* __restrict__ seg,
                                                                 const int* __restrict__ ncol_arr, const int* __restrict__ cp_off,
                                                                 const int* __restrict__ row1, const int* __restrict__ col1,
                                                                 const int* __restrict__ pad_off, int* __restrict__ colptr,
                                                                 int* __restrict__ rowidx, int* __restrict__ order) {
    extern __shared__ __align__(16) uint64_t wkeys[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= nmat) return;
    uint64_t* keys = wkeys + (size_t)warp * Pmax;
    const int z0 = seg[m], z = seg[m + 1] - z0, ncol = ncol_arr[m];
    const int P = pad_off[m + 1] - pad_off[m];
    int* cp = colptr + cp_off[m];
    for (int i = lane; i < P; i += 32) keys[i] = (i < z) ? pack_key(row1[z0 + i], col1[z0 + i], i) : ~0ull;
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int l = i ^ j;
                if (l > i) {
                    const uint64_t a = keys[i], b = keys[l];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncwarp();
        }
    }
    for (int i = lane; i < z; i += 32) {
        const uint64_t kk = keys[i];
        const int c = (int)(kk >> 42), r = (int)((kk >> 21) & 0x1fffff), cnt = (int)(kk & 0x1fffff);
        rowidx[z0 + i] = r;
        order[z0 + cnt] = i;
        const int cprev = (i == 0) ? -1 : (int)(keys[i - 1] >> 42);
        for (int j = cprev + 1; j <= c; j++) cp[j] = i;
    }
    const int clast = (z == 0) ? -1 : (int)(keys[z - 1] >> 42);
    for (int j = clast + 1 + lane; j <= ncol; j += 32) cp[j] = z;
}

// One warp per matrix with the keys in REGISTERS (R per lane, element e = r * 32 + lane): the compare-exchange partner of a
// bitonic stage is either another register of the same lane (distance >= 32: free) or the same register of another lane (one
// 64-bit shuffle), so the network needs no shared memory and no barrier.  The shared-memory network above costs ~3.3 k warp
// instructions per 115-entry matrix and bounded the kernel at 0.12 of the HBM roofline.  Same keys, same outputs.
// KEY = uint32_t when column, row and entry counter of the matrix fit 32 bits together (chosen per matrix, warp-uniform): half
// the instructions of the 64-bit network.
template <int R, typename KEY>
__device__ __forceinline__ void csc_reg_sort_emit(int lane, int z0, int z, int ncol, int cbits, int rbits, int nbits, const int* __restrict__ row1,
                                                  const int* __restrict__ col1, int* __restrict__ cp, int* __restrict__ rowidx,
                                                  int* __restrict__ order) {
    constexpr int P = 32 * R;
    (void)cbits;
    const KEY ALL = (KEY)~(KEY)0;
    KEY key[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        key[r] = (e < z) ? (KEY)(((KEY)(uint32_t)(col1[z0 + e] - 1) << (rbits + nbits)) | ((KEY)(uint32_t)(row1[z0 + e] - 1) << nbits) | (KEY)(uint32_t)e) : ALL;
    }
#pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int e = r * 32 + lane;
                const bool want_min = (((e & j) == 0) == ((e & k) == 0));
                if (j >= 32) {  // partner = register r ^ (j / 32) of this lane: handle the pair once, from its lower index
                    const int rp = r ^ (j >> 5);
                    if (rp > r) {
                        const KEY a = key[r], b = key[rp];
                        const bool swap = (a > b) == want_min;
                        key[r] = swap ? b : a; key[rp] = swap ? a : b;
                    }
                } else {
                    const KEY a = key[r], b = __shfl_xor_sync(0xffffffffu, a, j);
                    key[r] = want_min ? (a < b ? a : b) : (a > b ? a : b);
                }
            }
        }
    }
    // emit rowidx, order (inverse permutation) and colptr (exclusive counts) -- src/SpHbMat.cpp:255-264
    const KEY rmask = (KEY)(((KEY)1 << rbits) - 1), nmask = (KEY)(((KEY)1 << nbits) - 1);
    int clast = -1;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        const KEY kk = key[r];
        KEY prev = __shfl_up_sync(0xffffffffu, kk, 1);
        const KEY wrap = __shfl_sync(0xffffffffu, key[r > 0 ? r - 1 : 0], 31);
        if (lane == 0) prev = wrap;
        if (e < z) {
            const int c = (int)(kk >> (rbits + nbits)), rw = (int)((kk >> nbits) & rmask), cnt = (int)(kk & nmask);
            rowidx[z0 + e] = rw;
            order[z0 + cnt] = e;
            const int cprev = (e == 0) ? -1 : (int)(prev >> (rbits + nbits));
            for (int jj = cprev + 1; jj <= c; jj++) cp[jj] = e;
            if (e == z - 1) clast = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) clast = max(clast, __shfl_xor_sync(0xffffffffu, clast, o));
    for (int jj = clast + 1 + lane; jj <= ncol; jj += 32) cp[jj] = z;
}

template <int R>
__global__ void __launch_bounds__(256) csc_assemble_reg_kernel(int nmat, const int* __restrict__ seg, const int* __restrict__ ncol_arr,
                                                                const int* __restrict__ cp_off, const int* __restrict__ row1,
                                                                const int* __restrict__ col1, int* __restrict__ colptr,
                                                                int* __restrict__ rowidx, int* __restrict__ order) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= nmat) return;
    const int z0 = seg[m], z = seg[m + 1] - z0, ncol = ncol_arr[m];
    int* cp = colptr + cp_off[m];
    int rmax = 1;
    for (int e = lane; e < z; e += 32) rmax = max(rmax, row1[z0 + e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    const int cbits = 32 - __clz(max(ncol, 1)), rbits = 32 - __clz(rmax), nbits = 32 - __clz(max(z, 1));
    if (cbits + rbits + nbits <= 31) csc_reg_sort_emit<R, uint32_t>(lane, z0, z, ncol, cbits, rbits, nbits, row1, col1, cp, rowidx, order);
    else csc_reg_sort_emit<R, uint64_t>(lane, z0, z, ncol, 21, 21, 21, row1, col1, cp, rowidx, order);
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
    const bool small = total < (1LL << 31);
    for (; t < total; t += stride) {
        long long b;
        int k;
        if (small) { const unsigned tt = (unsigned)t, bb = tt / (unsigned)z_out; b = bb; k = (int)(tt - bb * (unsigned)z_out); }
        else { b = t / z_out; k = (int)(t - b * z_out); }
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
// A2: batched Vector operations (src/Vector.cpp:94-135, 174-184, 237-251; Utils oneNorm / infNorm src/Utils.cpp:65-83).
// Reductions run in the reference's index order per instance (bit-identical).  A CTA of 64 threads owns 64 instances: the
// rows move through shared memory in tiles of 32 columns (coalesced loads, conflict-free lane-per-row reads), so HBM sees
// every element once at full line width although each instance is summed sequentially.
// op: 0 getOneNorm, 1 getInfNorm, 2 times (dot product with y)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) vector_reduce_kernel(long long batch, int n, int op, const double* __restrict__ x,
                                                             const double* __restrict__ y, double* __restrict__ out) {
    __shared__ double tx[64][33];
    __shared__ double ty[64][33];
    const int tid = threadIdx.x, l = tid & 31, w = tid >> 5;
    for (long long b0 = (long long)blockIdx.x * 64; b0 < batch; b0 += (long long)gridDim.x * 64) {
        double acc = 0.0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            __syncthreads();
            for (int r = w; r < 64; r += 2) {  // warp w loads rows w, w+4, ...: 32 consecutive doubles of one instance
                const long long b = b0 + r;
                const int j = j0 + l;
                if (b < batch && j < n) { tx[r][l] = x[b * n + j]; if (op == 2) ty[r][l] = y[b * n + j]; }
            }
            __syncthreads();
            if (b0 + tid < batch) {
                const int jm = (n - j0 < 32) ? n - j0 : 32;
                for (int j = 0; j < jm; j++) {
                    const double v = tx[tid][j];
                    if (op == 0) acc = __dadd_rn(acc, fabs(v));
                    else if (op == 1) { const double a = (v < 0) ? -v : v; if (a > acc) acc = a; }
                    else acc = __dadd_rn(acc, __dmul_rn(v, ty[tid][j]));
                }
            }
        }
        if (b0 + tid < batch) out[b0 + tid] = acc;
    }
}
// elementwise: 0 add_vector (x += y), 1 subtract_vector (x -= y), 2 subtract_vector_to (x = y - x), 3 addNumber (x += alpha),
// 4 copy_vector (x = y), 5 scale (x *= alpha)
__global__ void __launch_bounds__(256) vector_elementwise_kernel(long long total, int op, double* __restrict__ x,
                                                                  const double* __restrict__ y, double alpha) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        const double v = x[t];
        double r;
        switch (op) {
        case 0: r = __dadd_rn(v, y[t]); break;
        case 1: r = __dsub_rn(v, y[t]); break;
        case 2: r = __dsub_rn(y[t], v); break;
        case 3: r = __dadd_rn(v, alpha); break;
        case 4: r = y[t]; break;
        default: r = __dmul_rn(v, alpha); break;
        }
        x[t] = r;
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

// Staged form: the values and the input vectors of a group of G consecutive instances are contiguous in HBM
// (instance-major layout), so the CTA copies them to shared memory with fully coalesced 16-byte loads, computes the
// G * nout outputs out of shared memory (one thread per output, same k order), and writes them with coalesced stores.
// Every input byte is fetched from HBM exactly once in full 128-byte lines; the shared pattern is staged once per CTA.
// Dynamic shared memory: G * (nnz + nin) doubles, then (nout + 1 + nnz [+ nnz]) ints.
__global__ void __launch_bounds__(256) spmv_staged_kernel(long long batch, int G, int nout, int nin, int nnz,
                                                           const int* __restrict__ ptr, const int* __restrict__ idx,
                                                           const int* __restrict__ perm, const double* __restrict__ val,
                                                           const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(16) double spm[];
    double* sval = spm;
    double* sx = sval + (size_t)G * nnz;
    int* sptr = reinterpret_cast<int*>(sx + (size_t)G * nin);
    int* sidx = sptr + nout + 1;
    int* sperm = sidx + nnz;
    for (int i = threadIdx.x; i <= nout; i += blockDim.x) sptr[i] = ptr[i];
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) { sidx[i] = idx[i]; if (perm) sperm[i] = perm[i]; }
    const long long ngroups = (batch + G - 1) / G;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long b0 = grp * G;
        const int g = (int)((batch - b0 < G) ? batch - b0 : G);
        __syncthreads();  // previous group consumed (and the pattern staged)
        {
            // b0 * nnz and b0 * nin are even (G is even), so both segments start 16-byte aligned
            const double2* v2 = reinterpret_cast<const double2*>(val + b0 * nnz);
            const double2* x2 = reinterpret_cast<const double2*>(x + b0 * nin);
            const int nv = g * nnz, nx = g * nin;
            for (int i = threadIdx.x; i < nv / 2; i += blockDim.x) reinterpret_cast<double2*>(sval)[i] = v2[i];
            if ((nv & 1) && threadIdx.x == 0) sval[nv - 1] = val[b0 * nnz + nv - 1];
            for (int i = threadIdx.x; i < nx / 2; i += blockDim.x) reinterpret_cast<double2*>(sx)[i] = x2[i];
            if ((nx & 1) && threadIdx.x == 0) sx[nx - 1] = x[b0 * nin + nx - 1];
        }
        __syncthreads();
        double* yg = y + b0 * nout;
        for (int t = threadIdx.x; t < g * nout; t += blockDim.x) {
            const int bl = t / nout, o = t - bl * nout;
            const double* vb = sval + (size_t)bl * nnz;
            const double* xb = sx + (size_t)bl * nin;
            double acc = 0.0;
            const int k1 = sptr[o + 1];
            for (int k = sptr[o]; k < k1; k++) {
                const int e = perm ? sperm[k] : k;
                acc = __dadd_rn(acc, __dmul_rn(vb[e], xb[sidx[k]]));
            }
            yg[t] = acc;
        }
    }
}

// TMA-staged, double-buffered form of the staged kernel: while the CTA computes group i out of one shared-memory stage, the
// TMA engine fills the other stage with group i+1 (two bulk copies: the G*nnz values and the G*nin inputs, both contiguous
// and 16-byte aligned because G is even).  No thread waits on a global load; every HBM line is fetched once.
// Dynamic shared memory: 2 stages x G*(nnz+nin) doubles, then the pattern as ints.
__global__ void __launch_bounds__(256) spmv_tma_kernel(long long batch, int G, int nout, int nin, int nnz,
                                                        const int* __restrict__ ptr, const int* __restrict__ idx,
                                                        const int* __restrict__ perm, const double* __restrict__ val,
                                                        const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(128) double spm[];
    __shared__ __align__(8) uint64_t mbar[2];
    const size_t stage_doubles = (size_t)G * (nnz + nin);
    int* sptr = reinterpret_cast<int*>(spm + 2 * stage_doubles);
    int* sidx = sptr + nout + 1;
    int* sperm = sidx + nnz;
    for (int i = threadIdx.x; i <= nout; i += blockDim.x) sptr[i] = ptr[i];
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) { sidx[i] = idx[i]; if (perm) sperm[i] = perm[i]; }
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ngroups = (batch + G - 1) / G;
    bool manual[2] = {false, false};
    uint32_t phase[2] = {0u, 0u};
    // fill stage s with group grp; returns true if the group had to be copied with ordinary loads (odd-sized tail)
    auto issue = [&](long long grp, int s) -> bool {
        const long long b0 = grp * G;
        const int g = (int)((batch - b0 < G) ? batch - b0 : G);
        double* sv = spm + (size_t)s * stage_doubles;
        double* sx = sv + (size_t)G * nnz;
        const uint32_t bv = (uint32_t)g * nnz * 8u, bx = (uint32_t)g * nin * 8u;
        if (((bv | bx) & 15u) == 0u) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&mbar[s], bv + bx);
                if (bv) bulk_g2s(sv, val + b0 * nnz, bv, &mbar[s]);
                if (bx) bulk_g2s(sx, x + b0 * nin, bx, &mbar[s]);
            }
            return false;
        }
        for (int i = threadIdx.x; i < g * nnz; i += blockDim.x) sv[i] = val[b0 * nnz + i];
        for (int i = threadIdx.x; i < g * nin; i += blockDim.x) sx[i] = x[b0 * nin + i];
        return true;
    };
    if ((long long)blockIdx.x < ngroups) manual[0] = issue(blockIdx.x, 0);
    __syncthreads();
    int it = 0;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, it++) {
        const int s = it & 1;
        const long long next = grp + gridDim.x;
        if (next < ngroups) manual[s ^ 1] = issue(next, s ^ 1);
        if (!manual[s]) {
            while (!mbar_try_wait(&mbar[s], phase[s])) {}
            phase[s] ^= 1u;
        }
        const long long b0 = grp * G;
        const int g = (int)((batch - b0 < G) ? batch - b0 : G);
        const double* sv = spm + (size_t)s * stage_doubles;
        const double* sx = sv + (size_t)G * nnz;
        double* yg = y + b0 * nout;
        for (int t = threadIdx.x; t < g * nout; t += blockDim.x) {
            const int bl = t / nout, o = t - bl * nout;
            const double* vb = sv + (size_t)bl * nnz;
            const double* xb = sx + (size_t)bl * nin;
            double acc = 0.0;
            const int k1 = sptr[o + 1];
            for (int k = sptr[o]; k < k1; k++) {
                const int e = perm ? sperm[k] : k;
                acc = __dadd_rn(acc, __dmul_rn(vb[e], xb[sidx[k]]));
            }
            yg[t] = acc;
        }
        __syncthreads();  // stage s is free again (and a manually copied next stage is complete)
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
    const bool small = total < (1LL << 31);  // 32-bit index arithmetic when it fits (64-bit division is emulated)
    for (; t < total; t += stride) {
        long long b;
        int i;
        if (small) { const unsigned tt = (unsigned)t, bb = tt / (unsigned)nV; b = bb; i = (int)(tt - bb * (unsigned)nV); }
        else { b = t / nV; i = (int)(t - b * nV); }
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
    const bool small = total < (1LL << 31);
    for (; t < total; t += stride) {
        long long b;
        int i;
        if (small) { const unsigned tt = (unsigned)t, bb = tt / (unsigned)nV; b = bb; i = (int)(tt - bb * (unsigned)nV); }
        else { b = t / nV; i = (int)(t - b * nV); }
        if (i < n) { if (grad) g[t] = grad[b * n + i]; }
        else if (rho) g[t] = rho[b];
    }
}

// ------------------------------------------------------------------------------------------
// C7 + C8 stand-alone: working-set translation and KKT residuals, one warp per instance.
// Lanes evaluate A x, A'y_c and H x (one lane per output entry, reference order); lane 0 then
// accumulates the four violation sums in the reference's index order, so the result is
// bit-identical with qpOASESInterface::test_optimality (src/qpOASESInterface.cpp:498-684).
// ------------------------------------------------------------------------------------------
struct KKTArgs {
    int batch, nV, nC, zA, zH, has_H;
    const int *Ap, *Ai, *Arp, *Aci, *Aperm, *Hp, *Hi;
    const double *Aval, *Hval, *g, *lb, *ub, *lbA, *ubA, *x, *y;
    const signed char *wsB, *wsC;
    int *WB, *WC;
    double* out;
};

// Shared memory per warp (doubles): av[zA] hv[zH] x[nV] y[nV+nC] g[nV] lb[nV] ub[nV] lbA[nC] ubA[nC] | Ax[nC] ATy[nV] Hx[nV]
// | terms[2nV+2nC]; then (bytes) sB[nV] sC[nC].  Everything an instance needs is copied from HBM once, coalesced; the lanes
// evaluate the three sparse products and then the individual violation terms in parallel; lanes 0..3 add the terms of the
// four sums one by one in the reference's index order (adding a term that the reference skips contributes +0.0, which
// leaves a non-negative partial sum unchanged), so the results stay bit-identical.
__host__ __device__ inline size_t kkt_warp_doubles(int nV, int nC, int zA, int zH) {
    size_t d = (size_t)zA + zH + nV + (nV + nC) + 3 * (size_t)nV + 2 * (size_t)nC + nC + 2 * (size_t)nV + 4 * (size_t)nV + 4 * (size_t)nC;
    return d + ((size_t)nV + nC + 7) / 8;
}

__global__ void __launch_bounds__(128) kkt_kernel(const KKTArgs A) {
    extern __shared__ __align__(16) double ksm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.batch) return;
    const int nV = A.nV, nC = A.nC, zA = A.zA, zH = A.has_H ? A.zH : 0;
    double* base = ksm + (size_t)warp * kkt_warp_doubles(nV, nC, A.zA, A.zH);
    double *av = base, *hv = av + zA, *x = hv + (A.has_H ? A.zH : 0) + (A.has_H ? 0 : A.zH), *y = x + nV, *g = y + nV + nC, *lb = g + nV, *ub = lb + nV;
    double *lbA = ub + nV, *ubA = lbA + nC, *Ax = ubA + nC, *ATy = Ax + nC, *Hx = ATy + nV;
    double *tP = Hx + nV, *tD = tP + 2 * nV + 2 * nC, *tC = tD + nV + nC;  // primal / dual / complementarity terms; stat terms reuse ATy
    signed char* sB = reinterpret_cast<signed char*>(tC + nV + nC);
    signed char* sC = sB + nV;
    {
        auto cp = [&](double* dst, const double* src, int cnt) { for (int i = lane; i < cnt; i += 32) dst[i] = src[i]; };
        cp(av, A.Aval + b * A.zA, zA);
        if (A.has_H) cp(hv, A.Hval + b * A.zH, zH);
        cp(x, A.x + b * nV, nV); cp(y, A.y + b * (nV + nC), nV + nC); cp(g, A.g + b * nV, nV);
        cp(lb, A.lb + b * nV, nV); cp(ub, A.ub + b * nV, nV); cp(lbA, A.lbA + b * nC, nC); cp(ubA, A.ubA + b * nC, nC);
        for (int i = lane; i < nV; i += 32) sB[i] = A.wsB[b * nV + i];
        for (int i = lane; i < nC; i += 32) sC[i] = A.wsC[b * nC + i];
    }
    __syncwarp();
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
        if (A.has_H)
            for (int e = A.Hp[c]; e < A.Hp[c + 1]; e++) h = __dadd_rn(h, __dmul_rn(hv[e], x[A.Hi[e]]));
        Hx[c] = h;
    }
    __syncwarp();
    const double SQRT_M_EPS = 1.0e-8;
    int* WB = A.WB ? A.WB + b * nV : nullptr;
    int* WC = A.WC ? A.WC + b * nC : nullptr;
#define WSB(i) ((sB[i] > 0) ? ((fabs(__dsub_rn(x[i], lb[i])) < SQRT_M_EPS) ? -99 : 1) \
                            : ((sB[i] < 0) ? ((fabs(__dsub_rn(x[i], ub[i])) < SQRT_M_EPS) ? -99 : -1) : 0))
#define WSC(i) ((sC[i] > 0) ? ((__dsub_rn(Ax[i], lbA[i]) < SQRT_M_EPS) ? -99 : 1) \
                            : ((sC[i] < 0) ? ((__dsub_rn(Ax[i], ubA[i]) < SQRT_M_EPS) ? -99 : -1) : 0))
    for (int i = lane; i < nV; i += 32) {
        tP[2 * i] = fmax(0.0, __dsub_rn(lb[i], x[i]));
        tP[2 * i + 1] = -fmin(0.0, __dsub_rn(ub[i], x[i]));
        const int W = WSB(i);
        const double yi = y[i];
        if (WB) WB[i] = W;
        tD[i] = (W == 0) ? fabs(yi) : ((W == -1) ? -fmin(0.0, yi) : ((W == 1) ? fmax(0.0, yi) : 0.0));
        tC[i] = (W == 0) ? fabs(yi) : ((W == -1) ? fabs(__dmul_rn(yi, __dsub_rn(x[i], lb[i]))) : ((W == 1) ? fabs(__dmul_rn(yi, __dsub_rn(ub[i], x[i]))) : 0.0));
        double gap = ATy[i];
        gap = __dadd_rn(gap, yi);
        gap = __dsub_rn(gap, g[i]);
        gap = __dsub_rn(gap, Hx[i]);
        ATy[i] = fabs(gap);  // stationarity term (each lane rewrites only its own entries)
    }
    for (int i = lane; i < nC; i += 32) {
        tP[2 * nV + 2 * i] = fmax(0.0, __dsub_rn(lbA[i], Ax[i]));
        tP[2 * nV + 2 * i + 1] = -fmin(0.0, __dsub_rn(ubA[i], Ax[i]));
        const int W = WSC(i);
        const double yi = y[nV + i];
        if (WC) WC[i] = W;
        tD[nV + i] = (W == 0) ? fabs(yi) : ((W == -1) ? -fmin(0.0, yi) : ((W == 1) ? fmax(0.0, yi) : 0.0));
        tC[nV + i] = (W == 0) ? fabs(yi) : ((W == -1) ? fabs(__dmul_rn(yi, __dsub_rn(Ax[i], lbA[i]))) : ((W == 1) ? fabs(__dmul_rn(yi, __dsub_rn(ubA[i], Ax[i]))) : 0.0));
    }
#undef WSB
#undef WSC
    __syncwarp();
    // lanes 0..3: one sequential sum each, in index order
    double s = 0.0;
    if (lane < 4) {
        const double* t = (lane == 0) ? tP : ((lane == 1) ? tD : ((lane == 2) ? ATy : tC));
        const int cnt = (lane == 0) ? 2 * nV + 2 * nC : ((lane == 2) ? nV : nV + nC);
        for (int i = 0; i < cnt; i++) s = __dadd_rn(s, t[i]);
    }
    const double primal = __shfl_sync(0xffffffffu, s, 0), dual = __shfl_sync(0xffffffffu, s, 1);
    const double stat = __shfl_sync(0xffffffffu, s, 2), cmpl = __shfl_sync(0xffffffffu, s, 3);
    if (lane == 0) {
        double* o = A.out + b * 5;
        o[0] = primal; o[1] = dual; o[2] = stat; o[3] = cmpl;
        o[4] = __dadd_rn(__dadd_rn(__dadd_rn(cmpl, stat), dual), primal);
    }
}

// TMA-staged form: a CTA of G warps (G even) works on groups of G consecutive instances, one warp per instance.  The nine
// FP64 arrays and the two working-set byte arrays of a group are contiguous in HBM; the TMA engine copies a group into a
// shared-memory stage (with two stages: the next group while the warps evaluate the current one).  The shared pattern is
// staged once per CTA.  Arithmetic and order of every sum as in kkt_kernel above (bit-identical results); the violation terms
// are written in place over staged inputs that are dead by then (lb/ub/lbA/ubA <- primal terms, g/Ax <- dual terms,
// x/y_c <- complementarity terms, A'y <- stationarity terms), so a warp needs only Ax, A'y, Hx as extra storage.  The byte
// arrays are fetched from the enclosing 16-byte aligned window (the library pads those allocations), `lead` bytes into it.
struct KKTLayout {
    int G, stages;
    int oAv, oHv, oX, oY, oG, oLb, oUb, oLbA, oUbA;  // stage offsets in doubles (each array: G * len)
    int oWsB, oWsC;                                    // stage offsets in bytes of the two byte windows
    int stage_bytes, work_doubles, pat_ints;           // per stage / per warp / per CTA
    int pAp, pAi, pArp, pAci, pAperm, pHp, pHi;        // pattern offsets in ints
};
__host__ __device__ inline KKTLayout kkt_layout(int G, int stages, int nV, int nC, int zA, int zH) {
    KKTLayout L;
    L.G = G; L.stages = stages;
    int o = 0;
    L.oAv = o; o += G * zA; L.oHv = o; o += G * zH; L.oX = o; o += G * nV; L.oY = o; o += G * (nV + nC);
    L.oG = o; o += G * nV; L.oLb = o; o += G * nV; L.oUb = o; o += G * nV; L.oLbA = o; o += G * nC; L.oUbA = o; o += G * nC;
    int bytes = o * 8;
    L.oWsB = bytes; bytes += ((G * nV + 15 + 15) / 16) * 16;
    L.oWsC = bytes; bytes += ((G * nC + 15 + 15) / 16) * 16;
    L.stage_bytes = (bytes + 127) / 128 * 128;
    L.work_doubles = nC + 2 * nV;
    int p = 0;
    L.pAp = p; p += nV + 1; L.pAi = p; p += zA; L.pArp = p; p += nC + 1; L.pAci = p; p += zA; L.pAperm = p; p += zA;
    L.pHp = p; p += nV + 1; L.pHi = p; p += zH;
    L.pat_ints = (p + 3) & ~3;
    return L;
}
__host__ __device__ inline size_t kkt_smem_bytes(const KKTLayout& L) {
    return (size_t)L.stages * L.stage_bytes + (size_t)L.G * L.work_doubles * 8 + (size_t)L.pat_ints * 4;
}

__global__ void __launch_bounds__(256) kkt_tma_kernel(const KKTArgs A, const KKTLayout L) {
    extern __shared__ __align__(128) unsigned char kraw[];
    __shared__ __align__(8) uint64_t mbar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = L.G, nV = A.nV, nC = A.nC, zA = A.zA, zH = A.has_H ? A.zH : 0;
    const int S = L.stages;
    double* work = reinterpret_cast<double*>(kraw + (size_t)S * L.stage_bytes) + (size_t)warp * L.work_doubles;
    int* pat = reinterpret_cast<int*>(reinterpret_cast<double*>(kraw + (size_t)S * L.stage_bytes) + (size_t)G * L.work_doubles);
    for (int i = threadIdx.x; i <= nV; i += blockDim.x) { pat[L.pAp + i] = A.Ap[i]; if (zH) pat[L.pHp + i] = A.Hp[i]; }
    for (int i = threadIdx.x; i <= nC; i += blockDim.x) pat[L.pArp + i] = A.Arp[i];
    for (int i = threadIdx.x; i < zA; i += blockDim.x) { pat[L.pAi + i] = A.Ai[i]; pat[L.pAci + i] = A.Aci[i]; pat[L.pAperm + i] = A.Aperm[i]; }
    for (int i = threadIdx.x; i < zH; i += blockDim.x) pat[L.pHi + i] = A.Hi[i];
    const int *Ap = pat + L.pAp, *Ai = pat + L.pAi, *Arp = pat + L.pArp, *Aci = pat + L.pAci, *Aperm = pat + L.pAperm, *Hp = pat + L.pHp, *Hi = pat + L.pHi;
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ngroups = ((long long)A.batch + G - 1) / G;
    bool manual[2] = {false, false};
    uint32_t phase[2] = {0u, 0u};
    int lead[2][2] = {{0, 0}, {0, 0}};  // [stage][B/C]: position of the group's first byte inside its aligned window
    auto issue = [&](long long grp, int s) -> bool {
        const long long b0 = grp * G;
        const int g = (int)((A.batch - b0 < G) ? A.batch - b0 : G);
        unsigned char* st = kraw + (size_t)s * L.stage_bytes;
        double* sd = reinterpret_cast<double*>(st);
        const unsigned char* gB = reinterpret_cast<const unsigned char*>(A.wsB) + b0 * nV;
        const unsigned char* gC = reinterpret_cast<const unsigned char*>(A.wsC) + b0 * nC;
        lead[s][0] = (int)((uintptr_t)gB & 15); lead[s][1] = (int)((uintptr_t)gC & 15);
        if ((g & 1) == 0) {  // even group: every FP64 segment is a multiple of 16 bytes and starts 16-byte aligned
            if (threadIdx.x == 0) {
                const uint32_t wB = (uint32_t)((lead[s][0] + g * nV + 15) / 16 * 16), wC = (uint32_t)((lead[s][1] + g * nC + 15) / 16 * 16);
                const uint32_t tot = 8u * (uint32_t)g * (uint32_t)(zA + zH + nV + (nV + nC) + 3 * nV + 2 * nC) + (nV ? wB : 0u) + (nC ? wC : 0u);
                // the stage was last written with ordinary stores (in-place terms): order them before the async-proxy writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar[s], tot);
                auto cp = [&](int off, const double* src, int len) { if (len) bulk_g2s(sd + off, src + b0 * len, 8u * (uint32_t)g * (uint32_t)len, &mbar[s]); };
                cp(L.oAv, A.Aval, zA); if (zH) cp(L.oHv, A.Hval, zH);
                cp(L.oX, A.x, nV); cp(L.oY, A.y, nV + nC); cp(L.oG, A.g, nV); cp(L.oLb, A.lb, nV); cp(L.oUb, A.ub, nV);
                cp(L.oLbA, A.lbA, nC); cp(L.oUbA, A.ubA, nC);
                if (nV) bulk_g2s(st + L.oWsB, gB - lead[s][0], wB, &mbar[s]);
                if (nC) bulk_g2s(st + L.oWsC, gC - lead[s][1], wC, &mbar[s]);
            }
            return false;
        }
        auto cpm = [&](int off, const double* src, int len) { for (int i = threadIdx.x; i < g * len; i += blockDim.x) sd[off + i] = src[b0 * len + i]; };
        cpm(L.oAv, A.Aval, zA); if (zH) cpm(L.oHv, A.Hval, zH);
        cpm(L.oX, A.x, nV); cpm(L.oY, A.y, nV + nC); cpm(L.oG, A.g, nV); cpm(L.oLb, A.lb, nV); cpm(L.oUb, A.ub, nV);
        cpm(L.oLbA, A.lbA, nC); cpm(L.oUbA, A.ubA, nC);
        for (int i = threadIdx.x; i < g * nV; i += blockDim.x) st[L.oWsB + lead[s][0] + i] = gB[i];
        for (int i = threadIdx.x; i < g * nC; i += blockDim.x) st[L.oWsC + lead[s][1] + i] = gC[i];
        return true;
    };
    if (S == 2 && (long long)blockIdx.x < ngroups) manual[0] = issue(blockIdx.x, 0);
    __syncthreads();
    int it = 0;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, it++) {
        const int s = (S == 2) ? (it & 1) : 0;
        if (S == 2) {
            const long long next = grp + gridDim.x;
            if (next < ngroups) manual[s ^ 1] = issue(next, s ^ 1);
        } else {
            manual[0] = issue(grp, 0);
            if (manual[0]) __syncthreads();
        }
        if (!manual[s]) {
            while (!mbar_try_wait(&mbar[s], phase[s])) {}
            phase[s] ^= 1u;
        }
        const long long b = grp * G + warp;
        if (b < A.batch) {
            unsigned char* st = kraw + (size_t)s * L.stage_bytes;
            double* sd = reinterpret_cast<double*>(st);
            const double *av = sd + L.oAv + (size_t)warp * zA, *hv = sd + L.oHv + (size_t)warp * zH;
            double *x = sd + L.oX + (size_t)warp * nV, *y = sd + L.oY + (size_t)warp * (nV + nC), *g = sd + L.oG + (size_t)warp * nV;
            double *lb = sd + L.oLb + (size_t)warp * nV, *ub = sd + L.oUb + (size_t)warp * nV;
            double *lbA = sd + L.oLbA + (size_t)warp * nC, *ubA = sd + L.oUbA + (size_t)warp * nC;
            const signed char* sB = reinterpret_cast<const signed char*>(st + L.oWsB + lead[s][0]) + (size_t)warp * nV;
            const signed char* sC = reinterpret_cast<const signed char*>(st + L.oWsC + lead[s][1]) + (size_t)warp * nC;
            double *Ax = work, *ATy = Ax + nC, *Hx = ATy + nV;
            for (int r = lane; r < nC; r += 32) {
                double acc = 0.0;
                const int k1 = Arp[r + 1];
                for (int k = Arp[r]; k < k1; k++) acc = __dadd_rn(acc, __dmul_rn(av[Aperm[k]], x[Aci[k]]));
                Ax[r] = acc;
            }
            for (int c = lane; c < nV; c += 32) {
                double acc = 0.0;
                const int e1 = Ap[c + 1];
                for (int e = Ap[c]; e < e1; e++) acc = __dadd_rn(acc, __dmul_rn(av[e], y[nV + Ai[e]]));
                ATy[c] = acc;
                double hh = 0.0;
                if (zH) {
                    const int h1 = Hp[c + 1];
                    for (int e = Hp[c]; e < h1; e++) hh = __dadd_rn(hh, __dmul_rn(hv[e], x[Hi[e]]));
                }
                Hx[c] = hh;
            }
            __syncwarp();
            const double SQRT_M_EPS = 1.0e-8;
            int* WB = A.WB ? A.WB + b * nV : nullptr;
            int* WC = A.WC ? A.WC + b * nC : nullptr;
            // terms, in place: lb[i], ub[i] <- primal terms; g[i] <- dual term; x[i] <- complementarity term; ATy[i] <- |gap|
            for (int i = lane; i < nV; i += 32) {
                const double xi = x[i], yi = y[i], lbi = lb[i], ubi = ub[i];
                const int sb = sB[i];
                const int W = (sb > 0) ? ((fabs(__dsub_rn(xi, lbi)) < SQRT_M_EPS) ? -99 : 1) : ((sb < 0) ? ((fabs(__dsub_rn(xi, ubi)) < SQRT_M_EPS) ? -99 : -1) : 0);
                if (WB) WB[i] = W;
                double gap = ATy[i];
                gap = __dadd_rn(gap, yi);
                gap = __dsub_rn(gap, g[i]);
                gap = __dsub_rn(gap, Hx[i]);
                ATy[i] = fabs(gap);
                lb[i] = fmax(0.0, __dsub_rn(lbi, xi));
                ub[i] = -fmin(0.0, __dsub_rn(ubi, xi));
                g[i] = (W == 0) ? fabs(yi) : ((W == -1) ? -fmin(0.0, yi) : ((W == 1) ? fmax(0.0, yi) : 0.0));
                x[i] = (W == 0) ? fabs(yi) : ((W == -1) ? fabs(__dmul_rn(yi, __dsub_rn(xi, lbi))) : ((W == 1) ? fabs(__dmul_rn(yi, __dsub_rn(ubi, xi))) : 0.0));
            }
            // constraints: lbA[i], ubA[i] <- primal terms; Ax[i] <- dual term; y[nV+i] <- complementarity term
            for (int i = lane; i < nC; i += 32) {
                const double ax = Ax[i], yi = y[nV + i], la = lbA[i], ua = ubA[i];
                const int sc = sC[i];
                const int W = (sc > 0) ? ((__dsub_rn(ax, la) < SQRT_M_EPS) ? -99 : 1) : ((sc < 0) ? ((__dsub_rn(ax, ua) < SQRT_M_EPS) ? -99 : -1) : 0);
                if (WC) WC[i] = W;
                lbA[i] = fmax(0.0, __dsub_rn(la, ax));
                ubA[i] = -fmin(0.0, __dsub_rn(ua, ax));
                Ax[i] = (W == 0) ? fabs(yi) : ((W == -1) ? -fmin(0.0, yi) : ((W == 1) ? fmax(0.0, yi) : 0.0));
                y[nV + i] = (W == 0) ? fabs(yi) : ((W == -1) ? fabs(__dmul_rn(yi, __dsub_rn(ax, la))) : ((W == 1) ? fabs(__dmul_rn(yi, __dsub_rn(ua, ax))) : 0.0));
            }
            __syncwarp();
            // lanes 0..3: one sequential sum each, in the reference's index order
            double acc = 0.0;
            if (lane == 0) {
                for (int i = 0; i < nV; i++) { acc = __dadd_rn(acc, lb[i]); acc = __dadd_rn(acc, ub[i]); }
                for (int i = 0; i < nC; i++) { acc = __dadd_rn(acc, lbA[i]); acc = __dadd_rn(acc, ubA[i]); }
            } else if (lane == 1) {
                for (int i = 0; i < nV; i++) acc = __dadd_rn(acc, g[i]);
                for (int i = 0; i < nC; i++) acc = __dadd_rn(acc, Ax[i]);
            } else if (lane == 2) {
                for (int i = 0; i < nV; i++) acc = __dadd_rn(acc, ATy[i]);
            } else if (lane == 3) {
                for (int i = 0; i < nV; i++) acc = __dadd_rn(acc, x[i]);
                for (int i = 0; i < nC; i++) acc = __dadd_rn(acc, y[nV + i]);
            }
            const double primal = __shfl_sync(0xffffffffu, acc, 0), dual = __shfl_sync(0xffffffffu, acc, 1);
            const double stat = __shfl_sync(0xffffffffu, acc, 2), cmpl = __shfl_sync(0xffffffffu, acc, 3);
            if (lane == 0) {
                double* o = A.out + b * 5;
                o[0] = primal; o[1] = dual; o[2] = stat; o[3] = cmpl;
                o[4] = __dadd_rn(__dadd_rn(__dadd_rn(cmpl, stat), dual), primal);
            }
        }
        __syncthreads();
    }
}

}  // namespace sqpb200
