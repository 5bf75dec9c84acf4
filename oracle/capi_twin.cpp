// oracle/capi_twin.cpp -- TEST INFRASTRUCTURE ONLY: a CPU twin of the part of the C ABI (include/sqpb200.h) that the two C++
// plugins call, backed by the CPU oracle (liboracle.so), batch = 1.  It exists for one purpose: to run the reference's own
// QPhandler.cpp -> CudaQPInterface.cpp / CudaQOREInterface.cpp host path end to end in a container without a GPU
// (oracle/_ref/qphandler_hs071_twin, tests/test_reference_qphandler.py), so that the plugins' glue and QPhandler's QORE / non-QORE
// branches are checked numerically on the CPU as well.  It is never linked into, loaded by or shipped with the product: the
// product library is restartsqp_b200/lib/libsqpb200.so, which has no CPU path.  Same state machine as the library's host side
// (src/qpOASESInterface.cpp:141-211, 817-833; handle_error :686-758), same as tests/oracle_backend.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/sqpb200.h"
extern "C" {
#include "oracle.h"
}

struct sqpb200_handle_s {
    int nV = 0, nC = 0, qptype = SQPB200_QP;
    sqpb200_options opt;
    std::string err;
    // column-compressed structure the solver works on, and the optional compressed-row view
    std::vector<int> Ap, Ai, Aorder, Hp, Hi, Horder, Jr, Jc, Hr, Hc;
    std::vector<double> Av, Hv;
    int zJ = 0, zHt = 0, Hsym = 1;
    bool A_set = false, H_set = false;
    struct Csr { bool set = false; std::vector<int> rp, ci, order, q2c; } qA, qH;
    std::vector<double> g, lb, ub, lbA, ubA, x, y, kkt;
    std::vector<int> wb, wc, WB, WC;
    double obj = 0.0;
    int status = 25, iters = 0;
    orc_qp* solver = nullptr;
    bool inited = false, first_solved = false, upd_A = false, upd_H = false;
    int old_ms = -1, new_ms = -1;
};

static void csr_view(int nrow, int ncol, const std::vector<int>& er, const std::vector<int>& ec, sqpb200_handle_s::Csr& q,
                     const std::vector<int>& Aorder) {
    // compressed-row arrays of the same entry list: the oracle's sort with the roles of row and column exchanged
    const int z = (int)er.size();
    q.rp.assign(nrow + 1, 0); q.ci.assign(z, 0); q.order.assign(z, 0);
    std::vector<double> dummy(z > 0 ? z : 1), zero(z > 0 ? z : 1, 0.0);
    orc_csc_from_entries(nrow, z, ec.data(), er.data(), zero.data(), q.rp.data(), q.ci.data(), dummy.data(), q.order.data());
    q.q2c.assign(z, 0);
    for (int i = 0; i < z; i++) q.q2c[q.order[i]] = Aorder[i];
    q.set = true;
    (void)ncol;
}

extern "C" {

void sqpb200_default_options(sqpb200_options* o) {
    memset(o, 0, sizeof *o);
    o->qp_maxiter = 1000; o->lp_maxiter = 100; o->enable_flipping = o->enable_ramping = o->enable_drift = 1; o->keep_state = 1;
}
const char* sqpb200_version(void) { return "sqpb200 CPU twin (test infrastructure, oracle-backed)"; }
int sqpb200_device_count(void) { return 1; }
const char* sqpb200_last_error(sqpb200_handle h) { return h ? h->err.c_str() : "null handle"; }

int sqpb200_create(int batch, int nV, int nC, int qptype, int, const sqpb200_options* opts, sqpb200_handle* out) {
    if (batch != 1 || nV <= 0 || nC < 0 || !out) return SQPB200_ERR_INVALID;
    sqpb200_handle h = new sqpb200_handle_s();
    h->nV = nV; h->nC = nC; h->qptype = qptype;
    if (opts) h->opt = *opts; else sqpb200_default_options(&h->opt);
    h->g.assign(nV, 0); h->lb.assign(nV, 0); h->ub.assign(nV, 0); h->lbA.assign(nC, 0); h->ubA.assign(nC, 0);
    h->x.assign(nV, 0); h->y.assign(nV + nC, 0); h->kkt.assign(5, 0);
    h->wb.assign(nV, 0); h->wc.assign(nC, 0); h->WB.assign(nV, 0); h->WC.assign(nC, 0);
    h->Ap.assign(nV + 1, 0); h->Hp.assign(nV + 1, 0);
    h->solver = orc_qp_create(nV, nC);
    *out = h;
    return 0;
}
int sqpb200_destroy(sqpb200_handle h) { if (!h) return SQPB200_ERR_INVALID; orc_qp_destroy(h->solver); delete h; return 0; }

static int structure_A(sqpb200_handle h, int zJ, const int* row1, const int* col1, int I_len, const int* I_irow, const int* I_jcol,
                       const int* I_size, const double* I_value, bool csr) {
    int z = zJ;
    for (int b = 0; b < I_len; b++) z += I_size[b];
    std::vector<int> er(z), ec(z);
    std::vector<double> ev(z), zero(zJ > 0 ? zJ : 1, 0.0);
    orc_expand_A(zJ, row1, col1, zero.data(), I_len, I_irow, I_jcol, I_size, I_value, er.data(), ec.data(), ev.data());
    h->Ai.assign(z, 0); h->Aorder.assign(z, 0); h->Av.assign(z, 0.0);
    orc_csc_from_entries(h->nV, z, er.data(), ec.data(), ev.data(), h->Ap.data(), h->Ai.data(), h->Av.data(), h->Aorder.data());
    h->zJ = zJ; h->A_set = true;
    h->qA.set = false;
    if (csr) csr_view(h->nC, h->nV, er, ec, h->qA, h->Aorder);
    return z;
}
int sqpb200_set_structure_A(sqpb200_handle h, int zJ, const int* r, const int* c, int L, const int* ir, const int* jc, const int* sz, const double* v) {
    return structure_A(h, zJ, r, c, L, ir, jc, sz, v, false);
}
int sqpb200_set_structure_A_csr(sqpb200_handle h, int zJ, const int* r, const int* c, int L, const int* ir, const int* jc, const int* sz, const double* v) {
    return structure_A(h, zJ, r, c, L, ir, jc, sz, v, true);
}
static int structure_H(sqpb200_handle h, int zH, const int* row1, const int* col1, int sym, bool csr) {
    std::vector<int> er(2 * zH + 1), ec(2 * zH + 1);
    std::vector<double> ev(2 * zH + 1), zero(zH > 0 ? zH : 1, 0.0);
    int z = orc_expand_H(zH, row1, col1, zero.data(), sym, er.data(), ec.data(), ev.data());
    er.resize(z); ec.resize(z);
    h->Hi.assign(z, 0); h->Horder.assign(z, 0); h->Hv.assign(z, 0.0);
    orc_csc_from_entries(h->nV, z, er.data(), ec.data(), ev.data(), h->Hp.data(), h->Hi.data(), h->Hv.data(), h->Horder.data());
    h->Hr.assign(row1, row1 + zH); h->Hc.assign(col1, col1 + zH); h->zHt = zH; h->Hsym = sym; h->H_set = true;
    h->qH.set = false;
    if (csr) csr_view(h->nV, h->nV, er, ec, h->qH, h->Horder);
    return z;
}
int sqpb200_set_structure_H(sqpb200_handle h, int zH, const int* r, const int* c, int sym) { return structure_H(h, zH, r, c, sym, false); }
int sqpb200_set_structure_H_csr(sqpb200_handle h, int zH, const int* r, const int* c, int sym) { return structure_H(h, zH, r, c, sym, true); }
// the data-constructor paths are not exercised by the QPhandler driver
int sqpb200_set_structure_csc(sqpb200_handle, int, int, const int*, const int*) { return SQPB200_ERR_STATE; }
int sqpb200_set_structure_csr(sqpb200_handle, int, int, const int*, const int*) { return SQPB200_ERR_STATE; }
int sqpb200_set_values_csc(sqpb200_handle, int, const double*, int, int) { return SQPB200_ERR_STATE; }
int sqpb200_set_values_csr(sqpb200_handle, int, const double*, int, int) { return SQPB200_ERR_STATE; }

int sqpb200_get_structure(sqpb200_handle h, int which, int* colptr, int* rowidx, int* order) {
    const auto& P = which == SQPB200_MAT_A ? h->Ap : h->Hp; const auto& I = which == SQPB200_MAT_A ? h->Ai : h->Hi;
    const auto& O = which == SQPB200_MAT_A ? h->Aorder : h->Horder;
    if (colptr) std::copy(P.begin(), P.end(), colptr);
    if (rowidx) std::copy(I.begin(), I.end(), rowidx);
    if (order) std::copy(O.begin(), O.end(), order);
    return 0;
}
int sqpb200_get_structure_csr(sqpb200_handle h, int which, int* rowptr, int* colidx, int* order) {
    const auto& q = which == SQPB200_MAT_A ? h->qA : h->qH;
    if (!q.set) return SQPB200_ERR_STATE;
    if (rowptr) std::copy(q.rp.begin(), q.rp.end(), rowptr);
    if (colidx) std::copy(q.ci.begin(), q.ci.end(), colidx);
    if (order) std::copy(q.order.begin(), q.order.end(), order);
    return 0;
}
int sqpb200_get_nnz(sqpb200_handle h, int which) { return (int)(which == SQPB200_MAT_A ? h->Ai.size() : h->Hi.size()); }
int sqpb200_set_values_A(sqpb200_handle h, const double* vals, int, int) {
    if (!h->A_set) return SQPB200_ERR_STATE;
    if (h->zJ > 0) orc_setmatval_A(h->zJ, h->Aorder.data(), vals, h->Av.data());
    if (h->first_solved) h->upd_A = true;
    return 0;
}
int sqpb200_set_values_H(sqpb200_handle h, const double* vals, int, int) {
    if (!h->H_set) return SQPB200_ERR_STATE;
    if (h->zHt > 0) orc_setmatval_H(h->zHt, h->Hr.data(), h->Hc.data(), h->Hsym, h->Horder.data(), vals, h->Hv.data());
    if (h->first_solved) h->upd_H = true;
    return 0;
}
int sqpb200_get_values_csc(sqpb200_handle h, int which, double* vals, int) {
    const auto& V = which == SQPB200_MAT_A ? h->Av : h->Hv;
    std::copy(V.begin(), V.end(), vals);
    return 0;
}
int sqpb200_get_values_csr(sqpb200_handle h, int which, double* vals, int) {
    const auto& q = which == SQPB200_MAT_A ? h->qA : h->qH; const auto& V = which == SQPB200_MAT_A ? h->Av : h->Hv;
    if (!q.set) return SQPB200_ERR_STATE;
    for (size_t e = 0; e < q.q2c.size(); e++) vals[e] = V[q.q2c[e]];
    return 0;
}
int sqpb200_set_vectors(sqpb200_handle h, int which, const double* vals, int offset, int count, int, int) {
    std::vector<double>* v = which == SQPB200_VEC_G ? &h->g : which == SQPB200_VEC_LB ? &h->lb : which == SQPB200_VEC_UB ? &h->ub :
                             which == SQPB200_VEC_LBA ? &h->lbA : &h->ubA;
    if (offset < 0 || offset + count > (int)v->size()) return SQPB200_ERR_INVALID;
    std::copy(vals, vals + count, v->begin() + offset);
    return 0;
}
int sqpb200_set_bounds_stacked(sqpb200_handle h, const double* lb, const double* ub, int, int) {
    if (lb) { std::copy(lb, lb + h->nV, h->lb.begin()); std::copy(lb + h->nV, lb + h->nV + h->nC, h->lbA.begin()); }
    if (ub) { std::copy(ub, ub + h->nV, h->ub.begin()); std::copy(ub + h->nV, ub + h->nV + h->nC, h->ubA.begin()); }
    return 0;
}

int sqpb200_solve(sqpb200_handle h, int mode, int, const unsigned char*) {
    const bool is_lp = mode == SQPB200_LP;
    orc_qp_options o;
    orc_qp_default_options(&o);
    o.max_iter = is_lp ? h->opt.lp_maxiter : h->opt.qp_maxiter;
    // init / hotstart decision of src/qpOASESInterface.cpp:141-211 with get_Matrix_change_status (:817-833)
    enum { COLD, FIXED, VARIED, REINIT } m = COLD;
    if (h->first_solved) {
        const bool varied = h->upd_A || h->upd_H;
        if (h->old_ms < 0) h->old_ms = varied ? 1 : 0;
        else { if (h->new_ms >= 0) h->old_ms = h->new_ms; h->new_ms = varied ? 1 : 0; }
        if (h->new_ms < 0) m = h->old_ms == 0 ? FIXED : VARIED;
        else if (h->new_ms == 0 && h->old_ms == 0) m = FIXED;
        else if (h->new_ms == 1 && h->old_ms == 1) m = VARIED;
        else { m = REINIT; h->new_ms = h->old_ms = -1; }
    }
    const int* Hp = is_lp ? nullptr : h->Hp.data(); const int* Hi = is_lp ? nullptr : h->Hi.data(); const double* Hv = is_lp ? nullptr : h->Hv.data();
    int st, its = 0;
    if (m == COLD || !h->inited)
        st = orc_qp_init(h->solver, &o, Hp, Hi, Hv, h->g.data(), h->Ap.data(), h->Ai.data(), h->Av.data(), h->lb.data(), h->ub.data(),
                         h->lbA.data(), h->ubA.data(), is_lp ? 1 : 0);
    else if (m == FIXED) st = orc_qp_hotstart(h->solver, &o, h->g.data(), h->lb.data(), h->ub.data(), h->lbA.data(), h->ubA.data());
    else if (m == VARIED) st = orc_qp_hotstart_matrices(h->solver, &o, Hv, h->Av.data(), h->g.data(), h->lb.data(), h->ub.data(), h->lbA.data(), h->ubA.data());
    else st = orc_qp_reinit(h->solver, &o, Hv, h->Av.data(), h->g.data(), h->lb.data(), h->ub.data(), h->lbA.data(), h->ubA.data());
    orc_qp_get_solution(h->solver, h->x.data(), h->y.data(), &h->obj, &its);
    const int st_first = st, its_first = its;
    if (st != 20) { int added = 0; st = orc_qp_handle_error(h->solver, &o, 0, &added); its += added; }
    h->inited = st == 20;
    int it2 = 0;
    orc_qp_get_solution(h->solver, h->x.data(), h->y.data(), &h->obj, &it2);
    if (st != 20) h->obj = 1.0e20;  // getObjVal of an unsolved problem
    h->status = st; h->iters = its;
    orc_qp_get_working_set(h->solver, h->wb.data(), h->wc.data());
    std::vector<double> Ax(h->nC > 0 ? h->nC : 1);
    orc_csc_times(h->nC, h->nV, h->Ap.data(), h->Ai.data(), h->Av.data(), h->x.data(), Ax.data());
    orc_translate_working_set(h->nV, h->nC, h->wb.data(), h->wc.data(), h->x.data(), Ax.data(), h->lb.data(), h->ub.data(), h->lbA.data(),
                              h->ubA.data(), h->WB.data(), h->WC.data());
    orc_kkt_residuals(h->nV, h->nC, h->Ap.data(), h->Ai.data(), h->Av.data(), Hp, Hi, Hv, h->g.data(), h->lb.data(), h->ub.data(), h->lbA.data(),
                      h->ubA.data(), h->x.data(), h->y.data(), h->WB.data(), h->WC.data(), h->kkt.data());
    if (getenv("SQPB200_TWIN_TRACE_DATA")) {  // one line per solve: every input vector and the solution as hex floats
        auto dump = [](const char* nm, const std::vector<double>& v) { fprintf(stderr, " %s", nm); for (double t : v) fprintf(stderr, " %a", t); };
        fprintf(stderr, "twin data:");
        dump("g", h->g); dump("lb", h->lb); dump("ub", h->ub); dump("lbA", h->lbA); dump("ubA", h->ubA); dump("Av", h->Av); dump("Hv", h->Hv); dump("x", h->x);
        fprintf(stderr, "\n");
    }
    if (getenv("SQPB200_TWIN_TRACE"))
        fprintf(stderr, "twin solve: first attempt %d / %d; mode %d status %d iters %d kkt %.3e %.3e %.3e %.3e | %.3e\n", st_first, its_first, (int)m, st, its, h->kkt[0], h->kkt[1], h->kkt[2],
                h->kkt[3], h->kkt[4]);
    h->upd_A = h->upd_H = false;
    h->first_solved = true;
    return 0;
}
int sqpb200_get_solution(sqpb200_handle h, double* x, double* y, double* obj, int* status, int* iters, int) {
    if (x) std::copy(h->x.begin(), h->x.end(), x);
    if (y) std::copy(h->y.begin(), h->y.end(), y);
    if (obj) *obj = h->obj;
    if (status) *status = h->status;
    if (iters) *iters = h->iters;
    return 0;
}
int sqpb200_get_solution_stacked(sqpb200_handle h, double* primal, double* dual, int* ws, int) {
    if (primal) {
        std::copy(h->x.begin(), h->x.end(), primal);
        if (h->nC > 0) orc_csc_times(h->nC, h->nV, h->Ap.data(), h->Ai.data(), h->Av.data(), h->x.data(), primal + h->nV);
    }
    if (dual) std::copy(h->y.begin(), h->y.end(), dual);
    if (ws) { for (int i = 0; i < h->nV; i++) ws[i] = -h->wb[i]; for (int i = 0; i < h->nC; i++) ws[h->nV + i] = -h->wc[i]; }
    return 0;
}
int sqpb200_get_working_set(sqpb200_handle h, int* wb, int* wc, int translated, int) {
    const auto& B = translated ? h->WB : h->wb; const auto& Cc = translated ? h->WC : h->wc;
    if (wb) std::copy(B.begin(), B.end(), wb);
    if (wc) std::copy(Cc.begin(), Cc.end(), wc);
    return 0;
}
int sqpb200_kkt_residuals(sqpb200_handle h, double* out, int) { std::copy(h->kkt.begin(), h->kkt.end(), out); return 0; }

}  // extern "C"
