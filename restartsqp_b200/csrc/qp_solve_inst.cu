// qp_solve_inst.cu -- one explicit instantiation of qp_solve_kernel<QP_CTA, QP_WPS> per object file
// (compiled once per (team size, CTA size), in parallel, by restartsqp_b200/build.py).
#include "qp_kernel.cuh"

#if !defined(QP_TEAM) || (QP_TEAM != 32 && QP_TEAM != 16 && QP_TEAM != 8)
#error "compile with -DQP_TEAM=<lanes per QP: 32, 16 or 8> -DQP_CTA=<threads per CTA: 32, 64 or 128>"
#endif

#ifndef QP_WPS
#define QP_WPS 16
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace sqpb200 {
cudaError_t CAT(CAT(CAT(CAT(CAT(launch_qp_solve_, QP_TEAM), _), QP_CTA), _w), QP_WPS)(const QPKernelArgs& a, int smem_bytes, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(qp_solve_kernel<QP_CTA, QP_WPS, QP_TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    const int teams = QP_CTA / QP_TEAM;
    int grid = (a.batch + teams - 1) / teams;
    if (a.rescue && grid > 296) grid = 296;  // rescue launch: a fixed small grid walks the (usually empty) list of overflowed instances
    qp_solve_kernel<QP_CTA, QP_WPS, QP_TEAM><<<grid, QP_CTA, smem_bytes, stream>>>(a);
    return cudaGetLastError();
}
}  // namespace sqpb200
