"""Plugin vocabulary of the reference (include/sqphot/Types.hpp, Options.hpp, Stats.hpp), mirrored
one-to-one so that host code and tests read like the reference's."""
import enum
from dataclasses import dataclass, field

import numpy as np

INF = 1.0e18          # include/sqphot/Utils.hpp:35
M_EPS = 1.0e-16       # :36
SQRT_M_EPS = 1.0e-8   # :37


class QPType(enum.IntEnum):  # Types.hpp:45-48
    LP = 1
    QP = 2


class Exitflag(enum.IntEnum):  # Types.hpp:51-73
    OPTIMAL = 0
    INVALID_NLP = -1
    CONVERGE_TO_NONOPTIMAL = 1
    EXCEED_MAX_ITER = 2
    PRED_REDUCTION_NEGATIVE = 3
    TRUST_REGION_TOO_SMALL = 4
    STEP_LARGER_THAN_TRUST_REGION = 5
    EXCEED_TIME_LIMITS = 6
    QP_UNCHANGED = 7  # not in Types.hpp: Algorithm::setupQP throws QP_UNCHANGED there (src/Algorithm.cpp:651-670); a batch cannot abort
    QP_OPTIMAL = 20
    QPERROR_INTERNAL_ERROR = 21
    QPERROR_INFEASIBLE = 22
    QPERROR_UNBOUNDED = 23
    QPERROR_EXCEED_MAX_ITER = 24
    QPERROR_NOTINITIALISED = 25
    QPERROR_PREPARINGAUXILIARYQP = 26
    QPERROR_AUXILIARYQPSOLVED = 27
    QPERROR_PERFORMINGHOMOTOPY = 28
    QPERROR_HOMOTOPYQPSOLVED = 29
    QPERROR_UNKNOWN = 30
    AUXINPUT_NOT_OPTIMAL = 99
    UNKNOWN = -99


class ActiveType(enum.IntEnum):  # Types.hpp:84-89
    ACTIVE_ABOVE = 1
    ACTIVE_BELOW = -1
    ACTIVE_BOTH_SIDE = -99
    INACTIVE = 0


class Solver(enum.IntEnum):  # Types.hpp:91-97 plus the one enumerator a maintainer adds
    QPOASES = 0
    QORE = 1
    GUROBI = 2
    CPLEX = 3
    SOLVER_UNDEFINED = 4
    CUDA_B200 = 5


@dataclass
class NLPInfo:  # Types.hpp:100-105
    nCon: int
    nVar: int
    nnz_jac_g: int = 0
    nnz_h_lag: int = 0


@dataclass
class IdentityInfo:  # Types.hpp:36-42
    irow: np.ndarray
    jcol: np.ndarray
    size: np.ndarray
    value: np.ndarray

    @property
    def length(self):
        return len(self.size)


@dataclass
class Options:  # Options::setToDefault, src/Options.cpp:19-57 (QPsolverChoice switched to this backend)
    iter_max: int = 1000
    time_max: float = 60.0
    printLevel: int = 2
    qpPrintLevel: int = 0
    QPsolverChoice: Solver = Solver.CUDA_B200
    LPsolverChoice: Solver = Solver.CUDA_B200
    second_order_correction: bool = False
    penalty_update: bool = True
    eta_c: float = 0.25
    eta_s: float = 1.0e-8
    eta_e: float = 0.75
    gamma_c: float = 0.5
    gamma_e: float = 2.0
    delta: float = 1.0
    delta_min: float = 1.0e-16
    delta_max: float = 1.0e8
    active_set_tol: float = 1.0e-5
    opt_stat_tol: float = 1.0e-4
    opt_compl_tol: float = 1.0e-4
    opt_dual_fea_tol: float = 1.0e-4
    opt_prim_fea_tol: float = 1.0e-4
    opt_second_tol: float = 1.0e-8
    tol: float = 1.0e-8
    penalty_update_tol: float = 1.0e-8
    rho: float = 1.0
    qp_maxiter: int = 1000
    increase_parm: float = 10.0
    rho_max: float = 1.0e6
    penalty_iter_max: int = 200
    eps1: float = 0.1
    eps1_change_parm: float = 0.1
    eps2: float = 1.0e-6
    EnablePertubation: bool = False
    lp_maxiter: int = 100


@dataclass
class Stats:  # include/sqphot/Stats.hpp:104-111 (qp_iter is per instance in batched mode)
    iter: int = 0
    qp_iter: np.ndarray = field(default_factory=lambda: np.zeros(1, dtype=np.int64))
    penalty_change_trial: int = 0
    penalty_change_Succ: int = 0
    penalty_change_Fail: int = 0
    soc_iter: int = 0

    def qp_iter_addValue(self, v):
        v = np.asarray(v, dtype=np.int64)
        if self.qp_iter.shape != v.shape:
            self.qp_iter = np.zeros(v.shape, dtype=np.int64)
        self.qp_iter = self.qp_iter + v


@dataclass
class SpTripletMat:
    """COO matrix, 1-based indices, optional symmetric-half storage (include/sqphot/SpTripletMat.hpp).
    MatVal may be [z] (shared by the batch) or [batch][z]."""
    RowIndex: np.ndarray
    ColIndex: np.ndarray
    MatVal: object
    RowNum: int
    ColNum: int
    isSymmetric: bool = False

    @property
    def EntryNum(self):
        return len(self.RowIndex)


class QP_NOT_OPTIMAL(Exception):  # include/sqphot/QPsolverInterface.hpp:26
    pass


class LP_NOT_OPTIMAL(Exception):  # :28
    pass


class QP_INTERNAL_ERROR(Exception):  # :30
    pass


class INVALID_WORKING_SET(Exception):  # :32
    pass
