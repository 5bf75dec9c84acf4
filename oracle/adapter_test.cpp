// oracle/adapter_test.cpp -- TEST INFRASTRUCTURE.  Drives the C++ drop-in plugin
// (restartsqp_b200/csrc/adapter/CudaQPInterface.cpp) exactly the way QPhandler drives a backend
// (src/QPhandler.cpp:167-201, 272-297, 310-334, 470-499) on the first QP subproblem of HS071, using the
// reference's own Vector / SpTripletMat / Options / Stats classes.  Built by oracle/Makefile (target `ref`) into
// oracle/_ref/adapter_hs071 where /root/reference is present; run by tests/test_gpu_adapter.py on the GPU box.
// With the argument `qore` it drives the QORE-layout plugin (CudaQOREInterface.cpp) instead.
// Output: one line per quantity, parsed by the test.
#include <cstdio>
#include <cmath>
#include <memory>
#include <cstring>
#include <CudaQPInterface.hpp>
#include <CudaQOREInterface.hpp>

using namespace SQPhotstart;

// `adapter_hs071 qore`: the same subproblem through the QORE-layout plugin, driven the way QPhandler's QORE branch drives a
// backend (src/QPhandler.cpp:225-260: constraint bounds at lb(nVar_QP + i), no set_lbA / set_ubA).
static int main_qore() {
    const int n = 4, m = 2, nV = n + 2 * m;
    NLPInfo info; info.nVar = n; info.nCon = m; info.nnz_jac_g = 8; info.nnz_h_lag = 10;
    auto options = std::make_shared<Options>();
    Ipopt::SmartPtr<Ipopt::Journalist> jnlst;
    std::shared_ptr<CudaQOREInterface> solver;
    try {
        solver = std::make_shared<CudaQOREInterface>(info, QP, options, jnlst);
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
    IdentityInfo I; int irow[2] = {1, 1}, jcol[2] = {n + 1, n + m + 1}, size[2] = {m, m}; double val[2] = {1.0, -1.0};
    I.length = 2; I.irow = irow; I.jcol = jcol; I.size = size; I.value = val;
    const double delta = 1.0, rho = 1.0, cl[2] = {25, 40}, cu[2] = {INF, 40};
    const double ck[2] = {xk[0] * xk[1] * xk[2] * xk[3], xk[0] * xk[0] + xk[1] * xk[1] + xk[2] * xk[2] + xk[3] * xk[3]};
    for (int i = 0; i < n; i++) { solver->set_lb(i, std::max(1.0 - xk[i], -delta)); solver->set_ub(i, std::min(5.0 - xk[i], delta)); }
    for (int i = 0; i < 2 * m; i++) solver->set_ub(n + i, INF);
    for (int i = 0; i < m; i++) { solver->set_lb(nV + i, cl[i] - ck[i]); solver->set_ub(nV + i, cu[i] - ck[i]); }
    solver->set_lbA(0, 123.0);  // a no-op in this layout
    const double grad[4] = {xk[3] * (2 * xk[0] + xk[1] + xk[2]), xk[0] * xk[3], xk[0] * xk[3] + 1, xk[0] * (xk[0] + xk[1] + xk[2])};
    for (int i = 0; i < nV; i++) solver->set_g(i, i < n ? grad[i] : rho);
    solver->set_A(J, I);
    solver->set_H(Hm);
    auto stats = std::make_shared<Stats>();
    try {
        solver->optimizeQP(stats);
    } catch (QP_NOT_OPTIMAL& e) {
        printf("not_optimal %d\n", (int)solver->get_status());
        return 3;
    }
    ActiveType Wb[8], Wc[2];
    bool ok = solver->test_optimality(Wc, Wb);
    printf("status %d\nqp_iter %d\nkkt_ok %d\nkkt_error %.17g\nobj %.17g\n", (int)solver->get_status(), stats->qp_iter, (int)ok,
           solver->get_optimality_status().KKT_error, solver->get_obj_value());
    printf("x"); for (int i = 0; i < nV + m; i++) printf(" %.17g", solver->get_optimal_solution()[i]); printf("\n");
    printf("y"); for (int i = 0; i < nV + m; i++) printf(" %.17g", solver->get_multipliers_bounds()[i]); printf("\n");
    printf("ws"); for (int i = 0; i < nV + m; i++) printf(" %d", solver->get_qore_working_set()[i]); printf("\n");
    printf("Wb"); for (int i = 0; i < nV; i++) printf(" %d", (int)Wb[i]); printf("\n");
    printf("Wc"); for (int i = 0; i < m; i++) printf(" %d", (int)Wc[i]); printf("\n");
    auto A = solver->getA();
    printf("A_rowptr"); for (int i = 0; i <= m; i++) printf(" %d", A->RowIndex(i)); printf("\n");
    printf("A_colidx"); for (int i = 0; i < A->EntryNum(); i++) printf(" %d", A->ColIndex(i)); printf("\n");
    printf("A_val"); for (int i = 0; i < A->EntryNum(); i++) printf(" %.17g", A->MatVal(i)); printf("\n");
    printf("lb"); for (int i = 0; i < nV + m; i++) printf(" %.17g", solver->getLb()->values(i)); printf("\n");
    int threw = 0;
    try { solver->getLbA(); } catch (INVALID_RETURN_TYPE_QORE_LAYOUT& e) { threw = 1; }
    printf("getLbA_threw %d\n", threw);
    // hot start with a smaller trust region (update_delta, src/QPhandler.cpp:559-564)
    for (int i = 0; i < n; i++) { solver->set_lb(i, std::max(1.0 - xk[i], -0.5)); solver->set_ub(i, std::min(5.0 - xk[i], 0.5)); }
    solver->optimizeQP(stats);
    printf("hot_status %d\nhot_qp_iter %d\n", (int)solver->get_status(), stats->qp_iter);
    printf("hot_x"); for (int i = 0; i < nV + m; i++) printf(" %.17g", solver->get_optimal_solution()[i]); printf("\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "qore")) return main_qore();
    const int n = 4, m = 2;
    NLPInfo info; info.nVar = n; info.nCon = m; info.nnz_jac_g = 8; info.nnz_h_lag = 10;
    auto options = std::make_shared<Options>();
    Ipopt::SmartPtr<Ipopt::Journalist> jnlst;
    std::shared_ptr<QPSolverInterface> solver;
    try {
        solver = std::make_shared<CudaQPInterface>(info, QP, options, jnlst);
    } catch (QP_INTERNAL_ERROR& e) {
        printf("create_failed %s\n", e.Message().c_str());
        return 2;
    }
    const double xk[4] = {1, 5, 5, 1};
    // Jacobian of (x1 x2 x3 x4, sum x_i^2), column-major triplets as AmplTNLP emits them
    auto J = std::make_shared<SpTripletMat>(8, m, n, false, true);
    const int jr[8] = {1, 2, 1, 2, 1, 2, 1, 2}, jc[8] = {1, 1, 2, 2, 3, 3, 4, 4};
    const double Jd[2][4] = {{xk[1] * xk[2] * xk[3], xk[0] * xk[2] * xk[3], xk[0] * xk[1] * xk[3], xk[0] * xk[1] * xk[2]},
                             {2 * xk[0], 2 * xk[1], 2 * xk[2], 2 * xk[3]}};
    for (int i = 0; i < 8; i++) { J->setRowIndex(i, jr[i]); J->setColIndex(i, jc[i]); J->setMatValAt(i, Jd[jr[i] - 1][jc[i] - 1]); }
    // Lagrangian Hessian (objective part + 2 I), upper triangle column by column, symmetric storage
    auto Hm = std::make_shared<SpTripletMat>(10, n, n, true, true);
    const int hr[10] = {1, 1, 2, 1, 2, 3, 1, 2, 3, 4}, hc[10] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4};
    double Hd[4][4] = {{2 * xk[3] + 2, xk[3], xk[3], 2 * xk[0] + xk[1] + xk[2]}, {0, 2, 0, xk[0]}, {0, 0, 2, xk[0]}, {0, 0, 0, 2}};
    for (int i = 0; i < 10; i++) { Hm->setRowIndex(i, hr[i]); Hm->setColIndex(i, hc[i]); Hm->setMatValAt(i, Hd[hr[i] - 1][hc[i] - 1]); }
    // I_info_A_ of QPhandler::QPhandler (src/QPhandler.cpp:41-51)
    IdentityInfo I; int irow[2] = {1, 1}, jcol[2] = {n + 1, n + m + 1}, size[2] = {m, m}; double val[2] = {1.0, -1.0};
    I.length = 2; I.irow = irow; I.jcol = jcol; I.size = size; I.value = val;
    // set_bounds (src/QPhandler.cpp:185-201), set_g (:287-292)
    const double delta = 1.0, rho = 1.0, cl[2] = {25, 40}, cu[2] = {INF, 40};
    const double ck[2] = {xk[0] * xk[1] * xk[2] * xk[3], xk[0] * xk[0] + xk[1] * xk[1] + xk[2] * xk[2] + xk[3] * xk[3]};
    for (int i = 0; i < m; i++) { solver->set_lbA(i, cl[i] - ck[i]); solver->set_ubA(i, cu[i] - ck[i]); }
    for (int i = 0; i < n; i++) { solver->set_lb(i, std::max(1.0 - xk[i], -delta)); solver->set_ub(i, std::min(5.0 - xk[i], delta)); }
    for (int i = 0; i < 2 * m; i++) solver->set_ub(n + i, INF);
    const double grad[4] = {xk[3] * (2 * xk[0] + xk[1] + xk[2]), xk[0] * xk[3], xk[0] * xk[3] + 1, xk[0] * (xk[0] + xk[1] + xk[2])};
    for (int i = 0; i < n + 2 * m; i++) solver->set_g(i, i < n ? grad[i] : rho);
    solver->set_A(J, I);
    solver->set_H(Hm);
    auto stats = std::make_shared<Stats>();
    try {
        solver->optimizeQP(stats);
    } catch (QP_NOT_OPTIMAL& e) {
        printf("not_optimal %d\n", (int)solver->get_status());
        return 3;
    }
    ActiveType Wb[8], Wc[2];
    bool ok = solver->test_optimality(Wc, Wb);
    printf("status %d\nqp_iter %d\nkkt_ok %d\nkkt_error %.17g\nobj %.17g\n", (int)solver->get_status(), stats->qp_iter, (int)ok,
           solver->get_optimality_status().KKT_error, solver->get_obj_value());
    printf("x"); for (int i = 0; i < n + 2 * m; i++) printf(" %.17g", solver->get_optimal_solution()[i]); printf("\n");
    printf("y"); for (int i = 0; i < n + 3 * m; i++) printf(" %.17g", solver->get_multipliers_bounds()[i]); printf("\n");
    printf("Wb"); for (int i = 0; i < n + 2 * m; i++) printf(" %d", (int)Wb[i]); printf("\n");
    printf("Wc"); for (int i = 0; i < m; i++) printf(" %d", (int)Wc[i]); printf("\n");
    auto A = solver->getA();
    printf("A_colptr"); for (int i = 0; i <= n + 2 * m; i++) printf(" %d", A->ColIndex(i)); printf("\n");
    printf("A_rowidx"); for (int i = 0; i < A->EntryNum(); i++) printf(" %d", A->RowIndex(i)); printf("\n");
    // hot start with a smaller trust region (update_delta, src/QPhandler.cpp:559-564)
    for (int i = 0; i < n; i++) { solver->set_lb(i, std::max(1.0 - xk[i], -0.5)); solver->set_ub(i, std::min(5.0 - xk[i], 0.5)); }
    solver->optimizeQP(stats);
    printf("hot_status %d\nhot_qp_iter %d\n", (int)solver->get_status(), stats->qp_iter);
    printf("hot_x"); for (int i = 0; i < n + 2 * m; i++) printf(" %.17g", solver->get_optimal_solution()[i]); printf("\n");
    return 0;
}
