// oracle/backend_pin_test.cpp -- TEST INFRASTRUCTURE.  The reference's REAL backend glue, src/qpOASESInterface.cpp (init / hotstart
// state machine :137-284 and :817-833, handle_error :686-758, get_status :330-357, get_working_set :835-895, test_optimality
// :498-684), running on a functional stand-in for qpOASES whose arithmetic is the oracle's (stubs_link/qpoases_over_oracle.cpp),
// side by side with the CUDA plugin CudaQPInterface on the CPU twin of the C ABI (capi_twin.cpp: the same oracle solver under the
// library's restated state machine).  Both receive the same random sequences of set_* / optimizeQP / optimizeLP calls -- cold
// start, hot starts with fixed and with new matrices, matrix-status flips, infeasible data -- and every observable is compared
// bit for bit: x, y, status, objective, Stats::qp_iter, the translated working sets, the five fields of OptimalityStatus, and
// whether optimizeQP threw.  What differs is printed (`mismatch ...`) and classified by tests/test_reference_backend.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include <sqphot/qpOASESInterface.hpp>
#include <CudaQPInterface.hpp>

using namespace SQPhotstart;

static unsigned long long rng_state = 88172645463325252ULL;
static double urand() {  // xorshift64*, uniform in (0, 1)
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return ((rng_state * 2685821657736338717ULL) >> 11) * (1.0 / 9007199254740992.0);
}
static double nrand() { double s = 0; for (int i = 0; i < 12; i++) s += urand(); return s - 6.0; }

struct Problem {
    int n, m;
    std::vector<int> jr, jc, hr, hc;
    std::vector<double> jv, hv, g, lb, ub, lbA, ubA;
};

static void new_matrices(Problem& P) {
    for (auto& v : P.jv) v = nrand();
    std::vector<double> S(P.n * P.n);
    for (auto& v : S) v = nrand();
    size_t k = 0;
    for (int c = 0; c < P.n; c++)
        for (int r = 0; r <= c; r++) {
            double h = (r == c) ? 0.5 : 0.0;
            for (int t = 0; t < P.n; t++) h += S[r * P.n + t] * S[c * P.n + t];
            P.hv[k++] = h;
        }
}
static void new_vectors(Problem& P, double delta, double rho, bool infeasible) {
    const int n = P.n, m = P.m, nV = n + 2 * m;
    P.g.assign(nV, rho); P.lb.assign(nV, 0.0); P.ub.assign(nV, INF); P.lbA.assign(m, 0.0); P.ubA.assign(m, 0.0);
    for (int i = 0; i < n; i++) { P.g[i] = nrand(); P.lb[i] = -delta; P.ub[i] = delta; }
    for (int i = 0; i < m; i++) {
        const double ck = nrand();
        const int kind = (int)(urand() * 3);
        P.lbA[i] = kind == 1 ? -INF : -ck;
        P.ubA[i] = kind == 2 ? -ck + 0.5 : (kind == 1 ? -ck : -ck);
        if (kind == 0) P.ubA[i] = P.lbA[i];
    }
    if (infeasible && m > 0) { P.lbA[0] = 1.0; P.ubA[0] = -1.0; }  // crossing bounds
}

template <class S>
static void load(S& s, const Problem& P, bool matrices, bool is_lp) {
    const int n = P.n, m = P.m, nV = n + 2 * m;
    for (int i = 0; i < nV; i++) { s.set_g(i, P.g[i]); s.set_lb(i, P.lb[i]); s.set_ub(i, P.ub[i]); }
    for (int i = 0; i < m; i++) { s.set_lbA(i, P.lbA[i]); s.set_ubA(i, P.ubA[i]); }
    if (!matrices) return;
    auto J = std::make_shared<SpTripletMat>((int)P.jr.size(), m, n, false, true);
    for (size_t k = 0; k < P.jr.size(); k++) { J->setRowIndex(k, P.jr[k]); J->setColIndex(k, P.jc[k]); J->setMatValAt(k, P.jv[k]); }
    IdentityInfo I; int irow[2] = {1, 1}, jcol[2] = {n + 1, n + m + 1}, size[2] = {m, m}; double val[2] = {1.0, -1.0};
    I.length = 2; I.irow = irow; I.jcol = jcol; I.size = size; I.value = val;
    s.set_A(J, I);
    if (is_lp) return;
    auto Hm = std::make_shared<SpTripletMat>((int)P.hr.size(), n, n, true, true);
    for (size_t k = 0; k < P.hr.size(); k++) { Hm->setRowIndex(k, P.hr[k]); Hm->setColIndex(k, P.hc[k]); Hm->setMatValAt(k, P.hv[k]); }
    s.set_H(Hm);
}

template <class S>
static int solve(S& s, std::shared_ptr<Stats> st, bool is_lp) {
    try { if (is_lp) s.optimizeLP(st); else s.optimizeQP(st); }
    catch (QP_NOT_OPTIMAL&) { return 1; }
    catch (LP_NOT_OPTIMAL&) { return 2; }
    return 0;
}

int main(int argc, char** argv) {
    const int problems = argc > 1 ? atoi(argv[1]) : 40;
    long mism = 0, steps = 0;
    for (int p = 0; p < problems; p++) {
        Problem P;
        P.n = 2 + (int)(urand() * 5); P.m = (int)(urand() * 5);
        const bool is_lp = (p % 4 == 3);
        const int n = P.n, m = P.m, nV = n + 2 * m;
        for (int c = 0; c < n; c++) for (int r = 0; r < m; r++) { P.jr.push_back(r + 1); P.jc.push_back(c + 1); }
        for (int c = 0; c < n; c++) for (int r = 0; r <= c; r++) { P.hr.push_back(r + 1); P.hc.push_back(c + 1); }
        P.jv.resize(P.jr.size()); P.hv.resize(P.hr.size());
        NLPInfo info; info.nVar = n; info.nCon = m; info.nnz_jac_g = (int)P.jr.size(); info.nnz_h_lag = (int)P.hr.size();
        auto options = std::make_shared<Options>();
        Ipopt::Journalist journalist;
        Ipopt::SmartPtr<Ipopt::Journalist> jnlst(&journalist);
        qpOASESInterface ref(info, is_lp ? LP : QP, options, jnlst);
        CudaQPInterface plug(info, is_lp ? LP : QP, options, jnlst);
        auto st_ref = std::make_shared<Stats>(), st_plug = std::make_shared<Stats>();
        // step kinds: 0 cold, 1 vectors only, 2 matrices + vectors, 3 infeasible vectors
        const int plan[12] = {0, 1, 1, 2, 2, 1, 1, 2, 3, 1, 2, 1};
        for (int step = 0; step < 12; step++, steps++) {
            const int kind = plan[step];
            if (kind == 0 || kind == 2) new_matrices(P);
            new_vectors(P, step % 3 == 2 ? 0.5 : 1.0, 1.0 + 9.0 * (step % 2), kind == 3);
            load(ref, P, kind == 0 || kind == 2, is_lp);
            load(plug, P, kind == 0 || kind == 2, is_lp);
            if (getenv("PIN_TRACE")) fprintf(stderr, "p %d step %d kind %d lp %d n %d m %d: loaded\n", p, step, kind, (int)is_lp, n, m);
            const int t1 = solve(ref, st_ref, is_lp);
            if (getenv("PIN_TRACE")) fprintf(stderr, "  ref solved (%d)\n", t1);
            const int t2 = solve(plug, st_plug, is_lp);
            if (getenv("PIN_TRACE")) fprintf(stderr, "  plugin solved (%d)\n", t2);
            auto bad = [&](const char* what) { printf("mismatch problem %d step %d kind %d %s: %s\n", p, step, kind, is_lp ? "LP" : "QP", what); mism++; };
            if (t1 != t2) bad("thrown");
            if ((int)ref.get_status() != (int)plug.get_status()) { char b[64]; snprintf(b, sizeof b, "status %d/%d", (int)ref.get_status(), (int)plug.get_status()); bad(b); }
            if (st_ref->qp_iter != st_plug->qp_iter) { char b[64]; snprintf(b, sizeof b, "qp_iter %d/%d", st_ref->qp_iter, st_plug->qp_iter); bad(b); st_plug->qp_iter = st_ref->qp_iter; }
            if (memcmp(ref.get_optimal_solution(), plug.get_optimal_solution(), sizeof(double) * nV)) bad("x");
            if (memcmp(ref.get_multipliers_bounds(), plug.get_multipliers_bounds(), sizeof(double) * (nV + m))) bad("y");
            if (!is_lp && ref.get_status() == QP_OPTIMAL && plug.get_status() == QP_OPTIMAL) {  // the reference never tests an LP (H_ is null there)
                if (ref.get_obj_value() != plug.get_obj_value()) bad("objective");
                std::vector<ActiveType> Wb1(nV), Wc1(m + 1), Wb2(nV), Wc2(m + 1);
                const bool o1 = ref.test_optimality(Wc1.data(), Wb1.data()), o2 = plug.test_optimality(Wc2.data(), Wb2.data());
                if (o1 != o2) bad("test_optimality");
                if (memcmp(Wb1.data(), Wb2.data(), sizeof(ActiveType) * nV) || memcmp(Wc1.data(), Wc2.data(), sizeof(ActiveType) * m)) bad("working set");
                const OptimalityStatus a = ref.get_optimality_status(), b = plug.get_optimality_status();
                if (a.primal_violation != b.primal_violation || a.dual_violation != b.dual_violation || a.stationarity_violation != b.stationarity_violation ||
                    a.compl_violation != b.compl_violation || a.KKT_error != b.KKT_error) bad("OptimalityStatus");
            }
        }
    }
    printf("summary problems %d steps %ld mismatches %ld\n", problems, steps, mism);
    return 0;
}
