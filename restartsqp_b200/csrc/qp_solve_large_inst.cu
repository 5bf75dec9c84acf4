// qp_solve_large_inst.cu -- instantiation of qp_solve_large_kernel<QP_CTA>: one QP per CTA, the slice (vectors and the
// nV x nV factors) in global memory.  Used when a QP does not fit the shared-memory resident warp kernel.
#include "qp_kernel.cuh"

#ifndef QP_CTA
#define QP_CTA 512
#endif

namespace sqpb200 {
cudaError_t launch_qp_solve_large(const QPKernelArgs& a, cudaStream_t stream) {
    qp_solve_large_kernel<QP_CTA><<<a.batch, QP_CTA, 0, stream>>>(a);
    return cudaGetLastError();
}
int qp_solve_large_threads() { return QP_CTA; }
}  // namespace sqpb200
