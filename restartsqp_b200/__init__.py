"""restartsqp_b200: B200-native batched QP-subproblem engine behind RestartSQP's QPSolverInterface.

The product is the CUDA library restartsqp_b200/lib/libsqpb200.so (C ABI: include/sqpb200.h); this
package holds only the host-side mirror of the reference's plugin interface for that path.
"""
from .sqp_types import (ActiveType, Exitflag, IdentityInfo, NLPInfo, Options, QPType, Solver, SpTripletMat, Stats,
                        QP_NOT_OPTIMAL, LP_NOT_OPTIMAL, QP_INTERNAL_ERROR, INVALID_WORKING_SET, INF)
from . import _capi as capi
from .qp_interface import CudaQPInterface
from .qore_layout import CudaQOREInterface
from .qp_handler import QPhandler
from .sqp_driver import BatchedSQP, HS071

__all__ = ["ActiveType", "Exitflag", "IdentityInfo", "NLPInfo", "Options", "QPType", "Solver", "SpTripletMat",
           "Stats", "QP_NOT_OPTIMAL", "LP_NOT_OPTIMAL", "QP_INTERNAL_ERROR", "INVALID_WORKING_SET", "INF", "capi",
           "CudaQPInterface", "CudaQOREInterface", "QPhandler", "BatchedSQP", "HS071"]
