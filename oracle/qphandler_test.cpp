// oracle/qphandler_test.cpp -- TEST INFRASTRUCTURE.  The reference's OWN QPhandler (src/QPhandler.cpp, unmodified except for
// integration/restartsqp_cuda_backend.patch) driving the CUDA plugins: what Algorithm::setupQP / solveQP do for the first QP
// subproblem of HS071 (src/Algorithm.cpp:645-697, 77) and for a trust-region update (update_delta), through
// QPhandler::set_bounds / set_g / set_H / set_A / solveQP / update_delta.  `qphandler_hs071` uses Solver CUDA_B200
// (CudaQPInterface, the patched default), `qphandler_hs071 qore` uses CUDA_B200_QORE_LAYOUT (CudaQOREInterface through
// QPhandler's QORE branches).  Built by oracle/Makefile (target `ref`) where /root/reference is present: the patched reference
// sources are compiled from a scratch copy together with the reference's own qpOASESInterface.cpp / QOREInterface.cpp, whose
// solver libraries are replaced by aborting stand-ins (oracle/stubs_link).  Output: one line per quantity, parsed by the test.
#include <cstdio>
#include <cstring>
#include <memory>
#include <sqphot/QPhandler.hpp>

using namespace SQPhotstart;

int main(int argc, char** argv) {
    const bool qore = argc > 1 && !strcmp(argv[1], "qore");
    const int n = 4, m = 2, nV = n + 2 * m;
    NLPInfo info; info.nVar = n; info.nCon = m; info.nnz_jac_g = 8; info.nnz_h_lag = 10;
    auto options = std::make_shared<Options>();  // patched default: QPsolverChoice = CUDA_B200
    if (qore) options->QPsolverChoice = options->LPsolverChoice = CUDA_B200_QORE_LAYOUT;
    Ipopt::Journalist journalist;
    Ipopt::SmartPtr<Ipopt::Journalist> jnlst(&journalist);
    std::shared_ptr<QPhandler> qp;
    try {
        qp = std::make_shared<QPhandler>(info, QP, jnlst, options);
    } catch (QP_INTERNAL_ERROR& e) {
        printf("create_failed %s\n", e.Message().c_str());
        return 2;
    }
    const double xk[4] = {1, 5, 5, 1};
    auto J = std::make_shared<SpTripletMat>(8, m, n, false, true);
    const int jr[8] = {1, 2, 1, 2, 1, 2, 1, 2}, jc[8] = {1, 1, 2, 2, 3, 3, 4, 4};
    const double Jd[2][4] = {{xk[1] * xk[2] * xk[3], xk[0] * xk[2] * xk[3], xk[0] * xk[1] * xk[3], xk[0] * xk[1] * xk[2]},
                             {2 * xk[0], 2 * xk[1], 2 * xk[2], 2 * xk[3]}};
    for (int i = 0; i < 8; i++) { J->setRowIndex(i, jr[i]); J->setColIndex(i, jc[i]); J->setMatValAt(i, Jd[jr[i] - 1][jc[i] - 1]); }
    auto Hm = std::make_shared<SpTripletMat>(10, n, n, true, true);
    const int hr[10] = {1, 1, 2, 1, 2, 3, 1, 2, 3, 4}, hc[10] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4};
    double Hd[4][4] = {{2 * xk[3] + 2, xk[3], xk[3], 2 * xk[0] + xk[1] + xk[2]}, {0, 2, 0, xk[0]}, {0, 0, 2, xk[0]}, {0, 0, 0, 2}};
    for (int i = 0; i < 10; i++) { Hm->setRowIndex(i, hr[i]); Hm->setColIndex(i, hc[i]); Hm->setMatValAt(i, Hd[hr[i] - 1][hc[i] - 1]); }
    const double xl[4] = {1, 1, 1, 1}, xu[4] = {5, 5, 5, 5}, cl[2] = {25, 40}, cu[2] = {INF, 40};
    const double ck[2] = {xk[0] * xk[1] * xk[2] * xk[3], xk[0] * xk[0] + xk[1] * xk[1] + xk[2] * xk[2] + xk[3] * xk[3]};
    const double grad[4] = {xk[3] * (2 * xk[0] + xk[1] + xk[2]), xk[0] * xk[3], xk[0] * xk[3] + 1, xk[0] * (xk[0] + xk[1] + xk[2])};
    auto v = [](int len, const double* a) { return std::make_shared<const Vector>(len, a); };
    auto x_l = v(n, xl), x_u = v(n, xu), x_k = v(n, xk), c_l = v(m, cl), c_u = v(m, cu), c_k = v(m, ck), g = v(n, grad);
    // Algorithm::setupQP, first iteration (src/Algorithm.cpp:672-682)
    qp->set_bounds(1.0, x_l, x_u, x_k, c_l, c_u, c_k);
    qp->set_g(g, 1.0);
    qp->set_H(Hm);
    qp->set_A(J);
    auto stats = std::make_shared<Stats>();
    try {
        qp->solveQP(stats, options);
    } catch (QP_NOT_OPTIMAL& e) {
        printf("not_optimal %d\n", (int)qp->get_status());
        return 3;
    }
    printf("status %d\nqp_iter %d\nkkt_error %.17g\nobj %.17g\ninfea_model %.17g\n", (int)qp->get_status(), stats->qp_iter,
           qp->get_QpOptimalStatus().KKT_error, qp->get_objective(), qp->get_infea_measure_model());
    printf("x"); for (int i = 0; i < nV; i++) printf(" %.17g", qp->get_optimal_solution()[i]); printf("\n");
    printf("yb"); for (int i = 0; i < nV; i++) printf(" %.17g", qp->get_multipliers_bounds()[i]); printf("\n");
    printf("yc"); for (int i = 0; i < m; i++) printf(" %.17g", qp->get_multipliers_constr()[i]); printf("\n");
    ActiveType Ac[2], Ab[8];
    qp->get_active_set(Ac, Ab);
    printf("Ab"); for (int i = 0; i < nV; i++) printf(" %d", (int)Ab[i]); printf("\n");
    printf("Ac"); for (int i = 0; i < m; i++) printf(" %d", (int)Ac[i]); printf("\n");
    // a rejected step shrinks the radius: QPhandler::update_delta (src/QPhandler.cpp:533-567), then a hot start
    qp->update_delta(0.5, x_l, x_u, x_k);
    qp->solveQP(stats, options);
    printf("hot_status %d\nhot_qp_iter %d\n", (int)qp->get_status(), stats->qp_iter);
    printf("hot_x"); for (int i = 0; i < nV; i++) printf(" %.17g", qp->get_optimal_solution()[i]); printf("\n");
    return 0;
}
