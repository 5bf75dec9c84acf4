// tools/peaks.cu -- microbenchmarks for the roofline denominators MEASURED_PEAKS.json does not hold
// (SURVEY.md 8d): FP64 FMA peak (dependent-chain-free DFMA stream) and shared-memory read bandwidth.
// Built into restartsqp_b200/lib/libsqpb200_peaks.so by __graft_entry__.build(); bench.py calls it once per run.
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) fp64_fma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) smem_read_kernel(double* out, int iters) {
    __shared__ double buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
    __syncthreads();
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; i++) {
        s0 += buf[idx]; s1 += buf[(idx + 256) & 4095]; s2 += buf[(idx + 512) & 4095]; s3 += buf[(idx + 768) & 4095];
        idx = (idx + 1024) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

extern "C" int peaks_measure(double* fp64_gflops, double* smem_gbs) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    double* out;
    if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_f = 1e30f, best_s = 1e30f, ms;
    const int it_f = 1 << 14, it_s = 1 << 13;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        fp64_fma_kernel<<<blocks, threads>>>(out, it_f, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_f) best_f = ms;
        cudaEventRecord(e0);
        smem_read_kernel<<<blocks, threads>>>(out, it_s);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_s) best_s = ms;
    }
    *fp64_gflops = (double)blocks * threads * it_f * 8.0 * 2.0 / (best_f * 1e-3) / 1e9;
    *smem_gbs = (double)blocks * threads * it_s * 4.0 * 8.0 / (best_s * 1e-3) / 1e9;
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
