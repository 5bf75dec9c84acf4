// sqp_outer.cu -- the per-instance steps of the Sl1QP outer loop (src/Algorithm.cpp) as device kernels, one thread per NLP
// instance: the iterate, the trust-region radius, the penalty parameter and all flags stay in HBM in SoA
// layout between the QP/LP solves and the NLP evaluations, which also run on the device; the host only sequences the
// launches and reads back a handful of counters per iteration (how many instances are still active, which Update_* flags are
// raised, how many need the penalty update).
//
// Every phase restates the corresponding reference routine with its sums in the reference's index order:
//   PH_FLAGS      loop condition of Algorithm::Optimize (:59-61) + the flag bookkeeping of setupQP (:645-697)
//   PH_AFTER_QP   QP status / KKT acceptance (QPhandler::solveQP :470-499, Algorithm :64-72), get_search_direction (:609),
//                 infea_measure_model (:592-594) and the entry test of update_penalty_parameter (:886-900)
//   PH_LP_AFTER   results of the feasibility LP (:900-912)
//   PH_PEN_CHECK  loop conditions of the two penalty while-loops (:914-972), rho_trial *= increase_parm
//   PH_PEN_AFTER  after a QP re-solve with rho_trial
//   PH_PEN_FINAL  accept / reject rho_trial (:975-1010)
//   PH_TRIAL      x_trial = x_k + p_k (:414-429)
//   PH_RATIO      cal_infea of the trial point (:577-602), ratio_test (:722-801), get_multipliers for accepted steps (:618-630)
//   PH_FINISH     new derivatives for accepted steps, iter++, check_optimality (:170-411), update_radius (:820-849)
//   PH_FINAL      EXCEED_MAX_ITER (:160-168)
//   PH_SOC_*      second_order_correction (:1140-1211, opt-in): corrected QP data, step p_k + s_k, second ratio test
#include "../../include/sqpb200.h"

#include <cuda_runtime.h>

namespace {

enum { EX_OPTIMAL = 0, EX_EXCEED_MAX_ITER = 2, EX_TRUST_REGION_TOO_SMALL = 4, EX_UNKNOWN = -99,
       EX_QP_UNCHANGED = SQPB200_EXIT_QP_UNCHANGED };
enum { CT_BOUNDED = 5, CT_EQUAL = -5, CT_BOUNDED_ABOVE = 9, CT_BOUNDED_BELOW = 1, CT_UNBOUNDED = 0 };
enum { UP_A = 1, UP_H = 2, UP_BOUNDS = 4, UP_DELTA = 8, UP_PENALTY = 16, UP_G = 32 };

__device__ __forceinline__ double cal_infea(const sqpb200_sqp_state& S, const double* c, int b) {
    double s = 0.0;
    for (int i = 0; i < S.m; i++) {
        const double ci = c[(size_t)b * S.m + i], cl = S.c_l[(size_t)b * S.m + i], cu = S.c_u[(size_t)b * S.m + i];
        const double below = (ci < cl) ? cl - ci : 0.0;
        const double above = (ci >= cl && ci > cu) ? ci - cu : 0.0;
        s = s + below + above;
    }
    return s;
}

__device__ __forceinline__ void get_multipliers(const sqpb200_sqp_state& S, int b) {
    const int nV = S.n + 2 * S.m;
    const double* y = S.qp_y + (size_t)b * (nV + S.m);
    for (int i = 0; i < S.m; i++) S.lam_c[(size_t)b * S.m + i] = y[nV + i];
    for (int i = 0; i < S.n; i++) S.lam_x[(size_t)b * S.n + i] = y[i];
}

// Algorithm::check_optimality for instance b (multipliers already refreshed); writes KKT_error, may set OPTIMAL
__device__ __forceinline__ void check_optimality(const sqpb200_sqp_state& S, int b, double* diff) {
    const int n = S.n, m = S.m;
    const double *mv = S.lam_x + (size_t)b * n, *mc = S.lam_c + (size_t)b * m;
    const double primal = S.infea[b];
    double dual = 0.0, compl_ = 0.0;
    for (int i = 0; i < n; i++) {
        const int t = S.bound_type[(size_t)b * n + i];
        dual = dual + (t == CT_BOUNDED_ABOVE ? fmax(mv[i], 0.0) : 0.0) + (t == CT_BOUNDED_BELOW ? -fmin(mv[i], 0.0) : 0.0);
    }
    for (int i = 0; i < m; i++) {
        const int t = S.cons_type[(size_t)b * m + i];
        dual = dual + (t == CT_BOUNDED_ABOVE ? fmax(mc[i], 0.0) : 0.0) + (t == CT_BOUNDED_BELOW ? -fmin(mc[i], 0.0) : 0.0);
    }
    for (int i = 0; i < m; i++) {
        const int t = S.cons_type[(size_t)b * m + i];
        const double ck = S.c_k[(size_t)b * m + i];
        compl_ = compl_ + (t == CT_BOUNDED_ABOVE ? fabs(mc[i] * (S.c_u[(size_t)b * m + i] - ck)) : 0.0)
                        + (t == CT_BOUNDED_BELOW ? fabs(mc[i] * (ck - S.c_l[(size_t)b * m + i])) : 0.0)
                        + (t == CT_UNBOUNDED ? fabs(mc[i]) : 0.0);
    }
    for (int i = 0; i < n; i++) {
        const int t = S.bound_type[(size_t)b * n + i];
        const double xk = S.x_k[(size_t)b * n + i];
        compl_ = compl_ + (t == CT_BOUNDED_ABOVE ? fabs(mv[i] * (S.x_u[(size_t)b * n + i] - xk)) : 0.0)
                        + (t == CT_BOUNDED_BELOW ? fabs(mv[i] * (xk - S.x_l[(size_t)b * n + i])) : 0.0)
                        + (t == CT_UNBOUNDED ? fabs(mv[i]) : 0.0);
    }
    // stationarity: || J'y_c + y_b - grad f ||_1, triplet SpMTV in storage order (src/SpTripletMat.cpp:311-323)
    for (int i = 0; i < n; i++) diff[i] = 0.0;
    const double* jv = S.jac + (size_t)b * S.zJ;
    for (int k = 0; k < S.zJ; k++) diff[S.J_col1[k] - 1] += jv[k] * mc[S.J_row1[k] - 1];
    double stat = 0.0;
    for (int i = 0; i < n; i++) {
        const double d = diff[i] + mv[i] - S.grad[(size_t)b * n + i];
        stat = stat + fabs(d);
    }
    S.kkt_err[b] = dual + primal + compl_ + stat;
    if (primal < S.opt_prim_fea_tol && dual < S.opt_dual_fea_tol && compl_ < S.opt_compl_tol && stat < S.opt_stat_tol) S.exitflag[b] = EX_OPTIMAL;
}

__global__ void sqp_phase_kernel(const __grid_constant__ sqpb200_sqp_state S, int phase) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    const int n = S.n, m = S.m, nV = n + 2 * m;
    switch (phase) {
    case SQPB200_PH_FLAGS: {
        bool a = S.iter[b] < S.iter_max && S.exitflag[b] == EX_UNKNOWN;
        // no Update_* flag raised since the instance's last solve: the reference throws QP_UNCHANGED (src/Algorithm.cpp:651-670),
        // which nothing catches; here the instance ends with its own exit flag instead of re-solving the same QP until iter_max
        if (a && S.iter[b] > 0 && S.upd[b] == 0) { S.exitflag[b] = EX_QP_UNCHANGED; a = false; }
        S.active[b] = a ? 1 : 0;
        if (a) {
            atomicAdd(&S.counters[0], 1);
            if (S.upd[b]) atomicOr(&S.counters[1], (int)S.upd[b]);
            // this instance's matrices change (it accepted a step): its backend sees Update_A / Update_H (:361-496)
            if (S.qp_inst && (S.upd[b] & (UP_A | UP_H)) && S.clear_flags) S.qp_inst[8 * (size_t)b + 1] = 1;
            if (S.clear_flags) S.upd[b] = 0;
        }
        break;
    }
    case SQPB200_PH_AFTER_QP: {
        S.rho_trial[b] = S.rho[b];
        if (!S.active[b]) { S.need[b] = 0; break; }
        S.qp_iter[b] += S.qp_iters[b];
        const int st = S.qp_status[b];
        const bool ok = (S.qp_kkt[(size_t)b * 5 + 4] <= 1.0e-6) && st == SQPB200_QP_OPTIMAL;
        if (!ok) {
            S.exitflag[b] = (st == SQPB200_QP_OPTIMAL) ? SQPB200_QPERROR_INTERNAL_ERROR : st;
            S.active[b] = 0; S.need[b] = 0;
            break;
        }
        const double* x = S.qp_x + (size_t)b * nV;
        for (int i = 0; i < n; i++) S.p_k[(size_t)b * n + i] = x[i];
        unsigned char need = 0;
        if (S.penalty_update) {
            double s = 0.0;
            for (int i = n; i < nV; i++) s = s + fabs(x[i]);
            S.infea_model[b] = s;
            need = s > S.penalty_update_tol;
            if (need) {
                atomicAdd(&S.counters[2], 1); S.infea_model_tmp[b] = s;
                if (S.lp_inst) S.lp_inst[8 * (size_t)b + 1] = 1;  // setupLP hands this instance's Jacobian to the LP backend (:700-704)
            }
        }
        S.need[b] = need;
        break;
    }
    case SQPB200_PH_LP_AFTER: {
        if (!S.need[b]) break;
        S.qp_iter[b] += S.lp_iters[b];
        const int st = S.lp_status[b];
        if (st != SQPB200_QP_OPTIMAL) { S.exitflag[b] = st; S.need[b] = 0; break; }  // LP_NOT_OPTIMAL leaves Optimize (:900-906)
        const double* x = S.lp_x + (size_t)b * nV;
        double s = 0.0;
        for (int i = n; i < nV; i++) s = s + fabs(x[i]);
        S.infea_infty[b] = s;
        S.feasible_lp[b] = s <= S.penalty_update_tol;
        break;
    }
    case SQPB200_PH_PEN_CHECK: {
        unsigned char go = 0;
        if (S.need[b] && S.exitflag[b] == EX_UNKNOWN) {
            const double rt = S.rho_trial[b];
            const bool cont_a = S.feasible_lp[b] && S.infea_model[b] > S.penalty_update_tol && rt < S.rho_max;
            const bool cont_b = !S.feasible_lp[b] && ((S.infea[b] - S.infea_model[b]) < S.eps1[b] * (S.infea[b] - S.infea_infty[b])) &&
                                S.pen_trial[b] < S.penalty_iter_max && rt < S.rho_max;
            if (cont_a || cont_b) {
                go = 1;
                S.rho_trial[b] = fmin(S.rho_max, rt * S.increase_parm);
                S.pen_trial[b] += 1;
                atomicAdd(&S.counters[3], 1);
            }
        }
        S.go[b] = go;
        break;
    }
    case SQPB200_PH_PEN_AFTER: {
        if (!S.go[b]) break;
        S.qp_iter[b] += S.qp_iters[b];
        const int st = S.qp_status[b];
        const bool ok = (S.qp_kkt[(size_t)b * 5 + 4] <= 1.0e-6) && st == SQPB200_QP_OPTIMAL;
        if (!ok) {
            // QP_NOT_OPTIMAL inside the penalty loop only leaves the loop (:932-935, :958-961): the acceptance test of
            // PH_PEN_FINAL then sees the objective of an unsolved QP (INFTY) and takes its failure branch, and the instance runs
            // the rest of the iteration (trial point, ratio test, iter++, check_optimality) before PH_FLAGS ends its solve
            S.exitflag[b] = (st == SQPB200_QP_OPTIMAL) ? SQPB200_QPERROR_INTERNAL_ERROR : st;
            S.need[b] = 2;
            break;
        }
        const double* x = S.qp_x + (size_t)b * nV;
        double s = 0.0;
        for (int i = n; i < nV; i++) s = s + fabs(x[i]);
        S.infea_model[b] = s;
        break;
    }
    case SQPB200_PH_PEN_FINAL: {
        if (!(S.need[b] && S.rho_trial[b] > S.rho[b] && (S.exitflag[b] == EX_UNKNOWN || S.need[b] == 2))) break;
        const double rt = S.rho_trial[b], qp_obj = S.qp_obj[b];
        const bool succ = rt * S.infea[b] - qp_obj >= S.eps2 * rt * (S.infea[b] - S.infea_model[b]);
        if (succ) {
            S.eps1[b] += (1 - S.eps1[b]) * S.eps1_change_parm;
            const double* x = S.qp_x + (size_t)b * nV;
            for (int i = 0; i < n; i++) S.p_k[(size_t)b * n + i] = x[i];
            S.rho[b] = rt;
        } else {
            S.infea_model[b] = S.infea_model_tmp[b];
            S.upd[b] |= UP_PENALTY;  // the backend still holds rho_trial in g: setupQP restores rho next iteration (:1003-1006)
        }
        break;
    }
    case SQPB200_PH_TRIAL: {
        if (S.active[b] && S.exitflag[b] != EX_UNKNOWN && S.need[b] != 2) S.active[b] = 0;
        if (!S.active[b]) break;
        double np_ = 0.0;  // ||p_k||_inf for update_radius (:822); an accepted second-order correction replaces it (SOC_RATIO)
        for (int i = 0; i < n; i++) {
            S.x_trial[(size_t)b * n + i] = S.x_k[(size_t)b * n + i] + S.p_k[(size_t)b * n + i];
            np_ = fmax(np_, fabs(S.p_k[(size_t)b * n + i]));
        }
        S.norm_p[b] = np_;
        break;
    }
    case SQPB200_PH_RATIO: {
        S.acc[b] = 0;
        if (!S.active[b]) break;
        const double infea_t = cal_infea(S, S.c_trial, b);
        S.infea_trial[b] = infea_t;
        const double P1_x = S.f_k[b] + S.rho[b] * S.infea[b];
        const double P1_t = S.f_trial[b] + S.rho[b] * infea_t;
        const double ared = P1_x - P1_t, pred = S.rho[b] * S.infea[b] - S.qp_obj[b];
        S.actual_red[b] = ared; S.pred_red[b] = pred;
        if (ared >= S.eta_s * pred && ared >= -S.tol) {
            S.acc[b] = 1;
            S.infea[b] = infea_t;
            S.f_k[b] = S.f_trial[b];
            for (int i = 0; i < n; i++) S.x_k[(size_t)b * n + i] = S.x_trial[(size_t)b * n + i];
            for (int i = 0; i < m; i++) S.c_k[(size_t)b * m + i] = S.c_trial[(size_t)b * m + i];
            get_multipliers(S, b);
            S.upd[b] |= UP_A | UP_H | UP_BOUNDS | UP_G;
            atomicAdd(&S.counters[4], 1);
        }
        // multipliers handed to the Hessian evaluation: -lambda (src/SQPTNLP.cpp:124-126)
        for (int i = 0; i < m; i++) S.neg_lam[(size_t)b * m + i] = -S.lam_c[(size_t)b * m + i];
        break;
    }
    case SQPB200_PH_FINISH: {
        if (!S.active[b]) break;
        if (S.acc[b]) {
            for (int i = 0; i < n; i++) S.grad[(size_t)b * n + i] = S.g_new[(size_t)b * n + i];
            for (int i = 0; i < S.zJ; i++) S.jac[(size_t)b * S.zJ + i] = S.j_new[(size_t)b * S.zJ + i];
            for (int i = 0; i < S.zH; i++) S.hess[(size_t)b * S.zH + i] = S.h_new[(size_t)b * S.zH + i];
        }
        S.iter[b] += 1;
        double* diff = S.scratch + (size_t)b * n;
        get_multipliers(S, b);
        check_optimality(S, b, diff);
        if (S.exitflag[b] != EX_UNKNOWN) break;
        // update_radius
        const double ared = S.actual_red[b], pred = S.pred_red[b];
        const bool shrink = ared < S.eta_c * pred;
        const double norm_p = S.norm_p[b];
        const bool grow = !shrink && ared > S.eta_e * pred && S.tol > fabs(S.delta[b] - norm_p);
        if (shrink) S.delta[b] = S.gamma_c * S.delta[b];
        if (grow) S.delta[b] = fmin(S.gamma_e * S.delta[b], S.delta_max);
        if (shrink || grow) S.upd[b] |= UP_DELTA;
        if (S.delta[b] < S.delta_min) {
            S.exitflag[b] = EX_TRUST_REGION_TOO_SMALL;
            check_optimality(S, b, diff);  // :149-152
        }
        break;
    }
    case SQPB200_PH_SOC_PREP: {
        // rejected steps: QP data of the corrected subproblem (gradient H_k p_k + g_k, bounds around the trial point); every
        // other instance keeps its current data in the mixed arrays
        const bool rj = S.active[b] && !S.acc[b] && S.exitflag[b] == EX_UNKNOWN;
        S.rej[b] = rj ? 1 : 0;
        double* sg = S.soc_g + (size_t)b * n;
        if (rj) {
            atomicAdd(&S.counters[5], 1);
            const double *hv = S.hess + (size_t)b * S.zH, *p = S.p_k + (size_t)b * n;
            for (int i = 0; i < n; i++) sg[i] = 0.0;
            for (int k = 0; k < S.zH; k++) {  // symmetric-half triplet product in storage order (src/SpTripletMat.cpp:237-258)
                const int i = S.H_row1[k] - 1, j = S.H_col1[k] - 1;
                sg[i] += hv[k] * p[j];
                if (i != j) sg[j] += hv[k] * p[i];
            }
            for (int i = 0; i < n; i++) { sg[i] = sg[i] + S.grad[(size_t)b * n + i]; S.p_tmp[(size_t)b * n + i] = p[i]; }
            S.qp_obj_tmp[b] = S.qp_obj[b];
        } else {
            for (int i = 0; i < n; i++) sg[i] = S.grad[(size_t)b * n + i];
        }
        for (int i = 0; i < n; i++) S.soc_x[(size_t)b * n + i] = rj ? S.x_trial[(size_t)b * n + i] : S.x_k[(size_t)b * n + i];
        for (int i = 0; i < m; i++) S.soc_c[(size_t)b * m + i] = rj ? S.c_trial[(size_t)b * m + i] : S.c_k[(size_t)b * m + i];
        break;
    }
    case SQPB200_PH_SOC_AFTER: {
        if (!S.rej[b]) break;
        S.qp_iter[b] += S.qp_iters[b];
        const int st = S.qp_status[b];
        const bool ok = (S.qp_kkt[(size_t)b * 5 + 4] <= 1.0e-6) && st == SQPB200_QP_OPTIMAL;
        if (!ok) {
            S.exitflag[b] = (st == SQPB200_QP_OPTIMAL) ? SQPB200_QPERROR_INTERNAL_ERROR : st;
            S.rej[b] = 2;  // left the correction with a solver failure: the saved step is restored in SOC_RATIO
            break;
        }
        const double* x = S.qp_x + (size_t)b * nV;
        S.qp_obj_soc[b] = S.qp_obj[b] + (S.qp_obj_tmp[b] - S.rho[b] * S.infea_model[b]);
        for (int i = 0; i < n; i++) {
            const double pk = S.p_k[(size_t)b * n + i] + x[i];
            S.p_k[(size_t)b * n + i] = pk;
            S.x_trial[(size_t)b * n + i] = S.x_k[(size_t)b * n + i] + pk;
        }
        break;
    }
    case SQPB200_PH_SOC_RATIO: {
        if (S.rej[b] == 1) {
            const double infea_t = cal_infea(S, S.c_trial, b);
            S.infea_trial[b] = infea_t;
            const double P1_x = S.f_k[b] + S.rho[b] * S.infea[b];
            const double P1_t = S.f_trial[b] + S.rho[b] * infea_t;
            // pred_reduction_ = rho * infea - get_obj_QP(): the raw objective of the SOC QP just solved (src/Algorithm.cpp:728);
            // qp_obj_soc mirrors the reference's qp_obj_ bookkeeping (:1197), which no test reads
            const double ared = P1_x - P1_t, pred = S.rho[b] * S.infea[b] - S.qp_obj[b];
            S.actual_red[b] = ared; S.pred_red[b] = pred;
            if (ared >= S.eta_s * pred && ared >= -S.tol) {
                S.acc[b] = 1;
                // update_radius reads p_k_->getInfNorm() (src/Algorithm.cpp:822): after an accepted correction that is the norm of the
                // corrected step p_k + s_k, which p_k holds since SOC_AFTER
                double np_ = 0.0;
                for (int i = 0; i < n; i++) np_ = fmax(np_, fabs(S.p_k[(size_t)b * n + i]));
                S.norm_p[b] = np_;
                S.infea[b] = infea_t;
                S.f_k[b] = S.f_trial[b];
                for (int i = 0; i < n; i++) S.x_k[(size_t)b * n + i] = S.x_trial[(size_t)b * n + i];
                for (int i = 0; i < m; i++) S.c_k[(size_t)b * m + i] = S.c_trial[(size_t)b * m + i];
                get_multipliers(S, b);
                S.upd[b] |= UP_A | UP_H | UP_BOUNDS | UP_G;
                for (int i = 0; i < m; i++) S.neg_lam[(size_t)b * m + i] = -S.lam_c[(size_t)b * m + i];
            } else {
                for (int i = 0; i < n; i++) S.p_k[(size_t)b * n + i] = S.p_tmp[(size_t)b * n + i];
            }
        } else if (S.rej[b] == 2) {
            for (int i = 0; i < n; i++) S.p_k[(size_t)b * n + i] = S.p_tmp[(size_t)b * n + i];
        }
        break;
    }
    case SQPB200_PH_INIT: {
        // Algorithm::initialization (src/Algorithm.cpp:438-472) per instance, after f, c and the derivatives have been evaluated
        // at the (shifted) start: infeasibility measure of the start, algorithm state back to the option values, backend state
        // machines back to "no QP solved yet".  Lets one DeviceBatchedSQP object (handles, buffers, compiled NLP) serve batch
        // after batch.
        S.infea[b] = cal_infea(S, S.c_k, b);
        S.delta[b] = S.delta0; S.rho[b] = S.rho0; S.eps1[b] = S.eps10;
        S.kkt_err[b] = __longlong_as_double(0x7ff0000000000000LL);
        S.exitflag[b] = EX_UNKNOWN; S.iter[b] = 0; S.pen_trial[b] = 0; S.qp_iter[b] = 0;
        S.active[b] = 0; S.need[b] = 0; S.go[b] = 0; S.acc[b] = 0; S.upd[b] = 0; S.feasible_lp[b] = 0; S.rej[b] = 0;
        for (int i = 0; i < n; i++) S.lam_x[(size_t)b * n + i] = 0.0;
        for (int i = 0; i < m; i++) S.neg_lam[(size_t)b * m + i] = -S.lam_c[(size_t)b * m + i];
        for (int k = 0; k < 2; k++) {
            signed char* st = k ? S.lp_inst : S.qp_inst;
            if (st) { st += 8 * (size_t)b; st[0] = 0; st[1] = 0; st[2] = -1; st[3] = -1; st[4] = 0; st[5] = 0; st[6] = 0; st[7] = 0; }
        }
        break;
    }
    case SQPB200_PH_FINAL: {
        if (S.iter[b] == S.iter_max && S.exitflag[b] == EX_UNKNOWN) S.exitflag[b] = EX_EXCEED_MAX_ITER;
        break;
    }
    }
}

}  // namespace

int sqpb200_sqp_phase(const sqpb200_sqp_state* st, int phase, int* counters_host, void* stream_) {
    if (!st || st->B <= 0) return SQPB200_ERR_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (phase == SQPB200_PH_FLAGS || phase == SQPB200_PH_AFTER_QP || phase == SQPB200_PH_PEN_CHECK || phase == SQPB200_PH_RATIO || phase == SQPB200_PH_SOC_PREP) {
        // counters: [0] active, [1] OR of Update_* flags, [2] need penalty update, [3] go, [4] accepted
        const int idx = phase == SQPB200_PH_FLAGS ? 0 : (phase == SQPB200_PH_AFTER_QP ? 2 : (phase == SQPB200_PH_PEN_CHECK ? 3 : (phase == SQPB200_PH_RATIO ? 4 : 5)));
        if (cudaMemsetAsync(st->counters + idx, 0, (phase == SQPB200_PH_FLAGS ? 2 : 1) * sizeof(int), stream) != cudaSuccess) return SQPB200_ERR_CUDA;
    }
    const int block = 128, grid = (st->B + block - 1) / block;
    sqp_phase_kernel<<<grid, block, 0, stream>>>(*st, phase);
    if (cudaGetLastError() != cudaSuccess) return SQPB200_ERR_CUDA;
    if (counters_host) {
        if (cudaMemcpyAsync(counters_host, st->counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return SQPB200_ERR_CUDA;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return SQPB200_ERR_CUDA;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Algorithm::Optimize (src/Algorithm.cpp:55-168) for the whole batch, sequenced from C++: the launch sequence of
// restartsqp_b200/sqp_device.py without an interpreter between the launches (at 10^4 instances the device work of an outer
// iteration is a few tens of microseconds, so the host side is what is timed).  The caller has set the structures of both
// handles (sqpb200_set_structure_A / _H) and evaluated the start point; `stream` must be the stream of both handles.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct OuterLoop {
    sqpb200_sqp_state* S;
    sqpb200_handle qp, lp;
    sqpb200_nlp nlp;
    cudaStream_t stream;
    int bounds_update_mode;
    int cnt[8];
    long long launches;
    int rc;
    bool ph(int phase, bool read = false) {
        rc = sqpb200_sqp_phase(S, phase, read ? cnt : nullptr, stream);
        launches++;
        return rc == 0;
    }
    bool solve(sqpb200_handle h, int type, const unsigned char* mask, signed char* inst) {
        rc = inst ? sqpb200_solve_per_instance(h, type, 0, mask, inst) : sqpb200_solve_device_mask(h, type, 0, mask);
        return rc == 0;
    }
    bool bounds(sqpb200_handle h, int mode, const double* xk, const double* ck) {
        rc = sqpb200_qphandler_bounds(h, mode, S->n, S->m, S->delta, S->x_l, S->x_u, xk, S->c_l, S->c_u, ck, SQPB200_LOC_DEVICE);
        return rc == 0;
    }
    bool grad(sqpb200_handle h, const double* g, const double* rho) {
        rc = sqpb200_qphandler_g(h, S->n, S->m, g, rho, SQPB200_LOC_DEVICE);
        return rc == 0;
    }
    // an empty Jacobian (m = 0) still raises Update_A like the reference's setter; the pointer is then never dereferenced
    bool valsA(sqpb200_handle h) { rc = sqpb200_set_values_A(h, S->zJ > 0 ? S->jac : S->x_k, SQPB200_LOC_DEVICE, 0); return rc == 0; }
    bool valsH(sqpb200_handle h) { rc = S->zH > 0 ? sqpb200_set_values_H(h, S->hess, SQPB200_LOC_DEVICE, 0) : 0; return rc == 0; }
    // setupQP, src/Algorithm.cpp:645-697.  The reference refreshes only what its Update_* flags name; here every refresh is
    // issued every iteration (rewriting unchanged data with itself), so that the host does not have to read the flags back
    // before it can queue the QP solve.  What the flags decide -- init / hotstart of each instance's backend -- is taken from
    // the per-instance flags on the device (PH_FLAGS), not from which setters were called.
    bool setupQP(int* first) {
        if (!valsA(qp) || !valsH(qp)) return false;
        if (*first) {
            if (!bounds(qp, 0, S->x_k, S->c_k)) return false;
            *first = 0;
            S->clear_flags = 1;
        } else if (!bounds(qp, bounds_update_mode, S->x_k, S->c_k)) return false;
        return grad(qp, S->grad, S->rho);
    }
    // the reference's own form (only what the Update_* bits name): used with handle-level init / hotstart decisions, where the
    // setters that were called decide the mode of the whole batch
    bool setupQP_bits(int bits, int* first) {
        if (*first) {
            if (!valsA(qp) || !valsH(qp) || !bounds(qp, 0, S->x_k, S->c_k) || !grad(qp, S->grad, S->rho)) return false;
            *first = 0;
            S->clear_flags = 1;
            return true;
        }
        if ((bits & 1) && !valsA(qp)) return false;
        if ((bits & 2) && !valsH(qp)) return false;
        if (bits & 4) { if (!bounds(qp, bounds_update_mode, S->x_k, S->c_k)) return false; }
        else if (bits & 8) { if (!bounds(qp, 2, S->x_k, nullptr)) return false; }
        if ((bits & 16) && !grad(qp, nullptr, S->rho)) return false;
        if ((bits & 32) && !grad(qp, S->grad, nullptr)) return false;
        return true;
    }
    // update_penalty_parameter, src/Algorithm.cpp:886-1028
    bool penalty() {
        if (!bounds(lp, 0, S->x_k, S->c_k) || !grad(lp, nullptr, S->rho) || !valsA(lp)) return false;  // setupLP :700-704
        if (!solve(lp, SQPB200_LP, S->need, S->lp_inst) || !ph(SQPB200_PH_LP_AFTER)) return false;
        for (;;) {
            if (!ph(SQPB200_PH_PEN_CHECK, true)) return false;
            if (cnt[3] == 0) break;
            if (!grad(qp, nullptr, S->rho_trial) || !solve(qp, SQPB200_QP, S->go, S->qp_inst) || !ph(SQPB200_PH_PEN_AFTER)) return false;
        }
        return ph(SQPB200_PH_PEN_FINAL);
    }
    // second_order_correction, src/Algorithm.cpp:1140-1211
    bool soc() {
        if (!ph(SQPB200_PH_SOC_PREP, true)) return false;
        if (cnt[5] == 0) return true;
        if (!grad(qp, S->soc_g, nullptr) || !bounds(qp, bounds_update_mode, S->soc_x, S->soc_c)) return false;
        if (!solve(qp, SQPB200_QP, S->rej, S->qp_inst) || !ph(SQPB200_PH_SOC_AFTER)) return false;
        rc = sqpb200_nlp_eval(nlp, 0, S->B, S->x_trial, nullptr, S->f_trial, S->c_trial, nullptr, nullptr, nullptr, SQPB200_LOC_DEVICE, stream);
        if (rc) return false;
        launches++;
        if (!ph(SQPB200_PH_SOC_RATIO)) return false;
        return grad(qp, S->grad, nullptr) && bounds(qp, bounds_update_mode, S->x_k, S->c_k);
    }
};
}  // namespace

int sqpb200_sqp_optimize(sqpb200_sqp_state* st, sqpb200_handle qp, sqpb200_handle lp, sqpb200_nlp nlp, int second_order_correction,
                         int refresh_ubA, int* first, double* f_tmp, double* c_tmp, long long* launches, void* stream) {
    if (!st || !qp || !lp || !nlp || !first || st->B <= 0) return SQPB200_ERR_INVALID;
    OuterLoop L{st, qp, lp, nlp, (cudaStream_t)stream, refresh_ubA ? 3 : 1, {0, 0, 0, 0, 0, 0, 0, 0}, 0, 0};
    for (;;) {
        // one read-back per outer iteration: the number of active instances (PH_FLAGS) and of instances that need the penalty
        // update (PH_AFTER_QP) come back together, after the QP solve has been queued behind the data refresh
        if (st->qp_inst) {
            if (!L.ph(SQPB200_PH_FLAGS)) return L.rc;
            if (!L.setupQP(first)) return L.rc;
        } else {
            if (!L.ph(SQPB200_PH_FLAGS, true)) return L.rc;
            if (L.cnt[0] == 0) break;
            if (!L.setupQP_bits(L.cnt[1], first)) return L.rc;
        }
        if (!L.solve(qp, SQPB200_QP, st->active, st->qp_inst)) return L.rc;
        if (!L.ph(SQPB200_PH_AFTER_QP, true)) return L.rc;
        if (L.cnt[0] == 0) break;  // nobody was active: the iteration above touched nothing
        if (L.cnt[2] > 0 && !L.penalty()) return L.rc;
        if (!L.ph(SQPB200_PH_TRIAL)) return L.rc;
        // get_trial_point_info :414-429
        L.rc = sqpb200_nlp_eval(nlp, 0, st->B, st->x_trial, nullptr, st->f_trial, st->c_trial, nullptr, nullptr, nullptr, SQPB200_LOC_DEVICE, stream);
        if (L.rc) return L.rc;
        if (!L.ph(SQPB200_PH_RATIO)) return L.rc;
        if (second_order_correction && !L.soc()) return L.rc;
        L.rc = sqpb200_nlp_eval(nlp, 1, st->B, st->x_k, st->neg_lam, f_tmp, c_tmp, st->g_new, st->j_new, st->h_new, SQPB200_LOC_DEVICE, stream);
        if (L.rc) return L.rc;
        L.launches += 2;
        if (!L.ph(SQPB200_PH_FINISH)) return L.rc;
    }
    if (!L.ph(SQPB200_PH_FINAL)) return L.rc;
    if (launches) *launches = L.launches;
    return 0;
}
