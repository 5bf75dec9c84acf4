/* oracle/oracle_batch.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Instance-parallel driver of the CPU oracle: one solver object per thread, POSIX threads over
 * independent instances (the reference is single-threaded per instance).  Used by
 * bench.py for the reported CPU baseline (`cpu_baseline`, `--impl reference`) and by tests that check
 * large batches; never by the product.
 * Timed region per QP, as BASELINE.md states: load data -> init (optimizeQP / optimizeLP) -> copy x, y.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    int B, nV, nC, Hv_stride, Av_stride, is_lp, max_iter;
    const int *Hp, *Hi, *Ap, *Ai;
    const double *Hv, *Av, *g, *lb, *ub, *lbA, *ubA;
    double *x, *y, *obj;
    int *status, *iters;
    atomic_int next;
} batch_job;

static void* batch_worker(void* arg) {
    batch_job* J = (batch_job*)arg;
    const int nV = J->nV, nC = J->nC, CHUNK = 8;
    orc_qp* q = orc_qp_create(nV, nC);
    orc_qp_options opt;
    orc_qp_default_options(&opt);
    opt.max_iter = J->max_iter;
    for (;;) {
        int b0 = atomic_fetch_add(&J->next, CHUNK);
        if (b0 >= J->B) break;
        int b1 = b0 + CHUNK < J->B ? b0 + CHUNK : J->B;
        for (int b = b0; b < b1; b++) {
            int st = orc_qp_init(q, &opt, J->is_lp ? NULL : J->Hp, J->Hi, J->is_lp ? NULL : J->Hv + (size_t)b * J->Hv_stride,
                                 J->g + (size_t)b * nV, J->Ap, J->Ai, J->Av + (size_t)b * J->Av_stride, J->lb + (size_t)b * nV,
                                 J->ub + (size_t)b * nV, J->lbA + (size_t)b * nC, J->ubA + (size_t)b * nC, J->is_lp);
            double o;
            int it, it1 = 0;
            orc_qp_get_solution(q, NULL, NULL, NULL, &it);
            if (st != ORC_QP_OPTIMAL) { st = orc_qp_handle_error(q, &opt, 0, &it1); it += it1; } /* optimizeQP -> handle_error */
            orc_qp_get_solution(q, J->x ? J->x + (size_t)b * nV : NULL, J->y ? J->y + (size_t)b * (nV + nC) : NULL, &o, NULL);
            if (J->obj) J->obj[b] = o;
            if (J->status) J->status[b] = st;
            if (J->iters) J->iters[b] = it;
        }
    }
    orc_qp_destroy(q);
    return NULL;
}

/* Cold-start solve of B instances sharing one pattern.  Value arrays have a per-instance stride
 * (0 = shared by all instances).  nthreads <= 0: all online cores.  Returns the number of threads used. */
int orc_qp_solve_batch(int B, int nV, int nC, const int* Hp, const int* Hi, const double* Hv, int Hv_stride,
                       const int* Ap, const int* Ai, const double* Av, int Av_stride, const double* g,
                       const double* lb, const double* ub, const double* lbA, const double* ubA, int is_lp,
                       int max_iter, double* x, double* y, double* obj, int* status, int* iters, int nthreads) {
    batch_job J;
    J.B = B; J.nV = nV; J.nC = nC; J.Hv_stride = Hv_stride; J.Av_stride = Av_stride; J.is_lp = is_lp; J.max_iter = max_iter;
    J.Hp = Hp; J.Hi = Hi; J.Ap = Ap; J.Ai = Ai; J.Hv = Hv; J.Av = Av; J.g = g; J.lb = lb; J.ub = ub; J.lbA = lbA; J.ubA = ubA;
    J.x = x; J.y = y; J.obj = obj; J.status = status; J.iters = iters;
    atomic_init(&J.next, 0);
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int t = 1; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, batch_worker, &J) == 0) started++;
    batch_worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    return started + 1;
}
