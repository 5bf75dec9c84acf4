// tma.cuh -- 1-D TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier) shared by the HBM-bound L0 kernels
// (l0_kernels.cuh) and the DMMA tile contraction of the one-QP-per-CTA solve kernel (qp_kernel.cuh).  SASS: UBLKCP.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqpb200 {
// ---- TMA helpers (1-D bulk asynchronous copy global -> shared, completion on an mbarrier) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return done != 0;
}

}  // namespace sqpb200
