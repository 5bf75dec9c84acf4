// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" shim around the *unmodified* reference L0 classes
// (SQPhotstart::SpTripletMat, SpHbMat, Vector; /root/reference/src/{Utils,Vector,
// SpTripletMat,SpHbMat}.cpp), compiled where those sources lie by oracle/Makefile
// into oracle/_ref/libref_l0.so.  It exists so that tests/ and tests/golden/make_golden.py
// can pin our own restatement (oracle/oracle_l0.c) and the CUDA kernels against the
// outputs of the real reference code: CSC index arrays, `order_` permutation, value
// scatter, SpMV / SpMTV and the 1-/inf-norms.
//
// Nothing in the product path may link or load this library; /root/reference does not
// exist on the GPU box, so this shim is only ever built in the dev container.
#include <memory>
#include <cstring>
#include <sqphot/SpHbMat.hpp>
#include <sqphot/SpTripletMat.hpp>
#include <sqphot/Vector.hpp>
#include <sqphot/Utils.hpp>

using namespace SQPhotstart;

static std::shared_ptr<SpTripletMat> make_triplet(int nrow, int ncol, int z, const int* row1,
                                                   const int* col1, const double* val,
                                                   bool symmetric) {
    auto t = std::make_shared<SpTripletMat>(z, nrow, ncol, symmetric, true);
    for (int i = 0; i < z; i++) {
        t->setRowIndex(i, row1[i]);
        t->setColIndex(i, col1[i]);
        t->setMatValAt(i, val ? val[i] : 0.0);
    }
    return t;
}

static void export_csc(const SpHbMat& m, int* colptr, int* rowidx, double* val, int* order) {
    for (int i = 0; i <= m.ColNum(); i++) colptr[i] = m.ColIndex(i);
    for (int i = 0; i < m.EntryNum(); i++) {
        rowidx[i] = m.RowIndex(i);
        if (val) val[i] = m.MatVal(i);
        order[i] = m.order(i);
    }
}

extern "C" {

// SpHbMat::setStructure(rhs, I_info)  (src/SpHbMat.cpp:196-268), CSC branch.
// Optionally followed by SpHbMat::setMatVal(rhs2, I_info) (src/SpHbMat.cpp:368-380) with new values.
int ref_assemble_A(int nrow, int ncol, int zJ, const int* row1, const int* col1, const double* val,
                   int I_len, int* I_irow, int* I_jcol, int* I_size, double* I_value,
                   const double* val_refresh, int* colptr, int* rowidx, double* out_val, int* order) {
    IdentityInfo info;
    info.length = I_len; info.irow = I_irow; info.jcol = I_jcol; info.size = I_size; info.value = I_value;
    int zI = 0;
    for (int i = 0; i < I_len; i++) zI += I_size[i];
    auto t = make_triplet(nrow, ncol, zJ, row1, col1, val, false);
    SpHbMat hb(zJ + zI, nrow, ncol, false);
    hb.setStructure(t, info);
    if (val_refresh) {
        auto t2 = make_triplet(nrow, ncol, zJ, row1, col1, val_refresh, false);
        hb.setMatVal(t2, info);
    }
    export_csc(hb, colptr, rowidx, out_val, order);
    return hb.EntryNum();
}

// SpHbMat::setStructure(rhs)  (src/SpHbMat.cpp:284-355), CSC branch, lazily allocated
// as in qpOASESInterface::allocate_memory (src/qpOASESInterface.cpp:122-124).
// Buffers must hold 2*zH entries.  Returns the full-symmetric nnz.
int ref_assemble_H(int n, int zH, const int* row1, const int* col1, const double* val, int symmetric,
                   const double* val_refresh, int* colptr, int* rowidx, double* out_val, int* order) {
    auto t = make_triplet(n, n, zH, row1, col1, val, symmetric != 0);
    SpHbMat hb(n, n, false);
    hb.setStructure(t);
    if (val_refresh) {
        auto t2 = make_triplet(n, n, zH, row1, col1, val_refresh, symmetric != 0);
        hb.setMatVal(t2);
    }
    export_csc(hb, colptr, rowidx, out_val, order);
    return hb.EntryNum();
}

static std::shared_ptr<SpHbMat> csc_from_arrays(int nrow, int ncol, int nnz, const int* colptr,
                                                 const int* rowidx, const double* val) {
    auto m = std::make_shared<SpHbMat>(nnz, nrow, ncol, false);
    for (int i = 0; i <= ncol; i++) m->setColIndexAt(i, colptr[i]);
    for (int i = 0; i < nnz; i++) {
        m->setRowIndexAt(i, rowidx[i]);
        m->setMatValAt(i, val[i]);
    }
    return m;
}

// SpHbMat::times (src/SpHbMat.cpp:698-737), CSC branch.
void ref_csc_times(int nrow, int ncol, int nnz, const int* colptr, const int* rowidx,
                   const double* val, const double* x, double* y) {
    auto m = csc_from_arrays(nrow, ncol, nnz, colptr, rowidx, val);
    auto p = std::make_shared<Vector>(ncol, x);
    auto r = std::make_shared<Vector>(nrow);
    if (nnz > 0) m->times(p, r);
    std::memcpy(y, r->values(), sizeof(double) * nrow);
}

// SpHbMat::transposed_times(const double*, double*) (include/sqphot/SpHbMat.hpp:140-177).
void ref_csc_transposed_times(int nrow, int ncol, int nnz, const int* colptr, const int* rowidx,
                              const double* val, const double* x, double* y) {
    auto m = csc_from_arrays(nrow, ncol, nnz, colptr, rowidx, val);
    if (nnz > 0) m->transposed_times(x, y);
    else for (int i = 0; i < ncol; i++) y[i] = 0.0;
}

// SpTripletMat::times / transposed_times (src/SpTripletMat.cpp:237-258, 311-323).
void ref_triplet_times(int nrow, int ncol, int z, const int* row1, const int* col1,
                       const double* val, int symmetric, int transpose, const double* x, double* y) {
    auto t = make_triplet(nrow, ncol, z, row1, col1, val, symmetric != 0);
    int nin = transpose ? nrow : ncol, nout = transpose ? ncol : nrow;
    auto p = std::make_shared<Vector>(nin, x);
    auto r = std::make_shared<Vector>(nout);
    if (transpose) t->transposed_times(p, r); else t->times(p, r);
    std::memcpy(y, r->values(), sizeof(double) * nout);
}

// Utils oneNorm / infNorm (src/Utils.cpp:65-83) and Vector::getOneNorm/getInfNorm.
double ref_one_norm(const double* x, int n) { return oneNorm(x, n); }
double ref_inf_norm(const double* x, int n) { return infNorm(x, n); }
double ref_vector_one_norm(const double* x, int n) { Vector v(n, x); return v.getOneNorm(); }
double ref_vector_inf_norm(const double* x, int n) { Vector v(n, x); return v.getInfNorm(); }
double ref_const_INF(void) { return INF; }
double ref_const_sqrt_m_eps(void) { return sqrt_m_eps; }

// ---- compressed-row ("QORE layout") branch of the same routines: what QOREInterface::set_A / set_H build
// (src/QOREInterface.cpp:643-659 on SpHbMat(..., isCompressedRow = true), include/sqphot/QOREInterface.hpp) and what
// QPSetData receives (src/QOREInterface.cpp:89-90: A_->RowIndex() = row pointers, A_->ColIndex() = column of each entry).
static void export_csr(const SpHbMat& m, int* rowptr, int* colidx, double* val, int* order) {
    for (int i = 0; i <= m.RowNum(); i++) rowptr[i] = m.RowIndex(i);
    for (int i = 0; i < m.EntryNum(); i++) {
        colidx[i] = m.ColIndex(i);
        if (val) val[i] = m.MatVal(i);
        order[i] = m.order(i);
    }
}

// SpHbMat::setStructure(rhs, I_info) (src/SpHbMat.cpp:196-268), compressed-row branch (:238-250), optionally followed by
// SpHbMat::setMatVal(rhs2, I_info).
int ref_assemble_A_csr(int nrow, int ncol, int zJ, const int* row1, const int* col1, const double* val,
                       int I_len, int* I_irow, int* I_jcol, int* I_size, double* I_value,
                       const double* val_refresh, int* rowptr, int* colidx, double* out_val, int* order) {
    IdentityInfo info;
    info.length = I_len; info.irow = I_irow; info.jcol = I_jcol; info.size = I_size; info.value = I_value;
    int zI = 0;
    for (int i = 0; i < I_len; i++) zI += I_size[i];
    auto t = make_triplet(nrow, ncol, zJ, row1, col1, val, false);
    SpHbMat hb(zJ + zI, nrow, ncol, true);
    hb.setStructure(t, info);
    if (val_refresh) {
        auto t2 = make_triplet(nrow, ncol, zJ, row1, col1, val_refresh, false);
        hb.setMatVal(t2, info);
    }
    export_csr(hb, rowptr, colidx, out_val, order);
    return hb.EntryNum();
}

// SpHbMat::setStructure(rhs) (src/SpHbMat.cpp:284-355), compressed-row branch (:324-337), lazily allocated as in
// QOREInterface::allocate_memory (src/QOREInterface.cpp:212).
int ref_assemble_H_csr(int n, int zH, const int* row1, const int* col1, const double* val, int symmetric,
                       const double* val_refresh, int* rowptr, int* colidx, double* out_val, int* order) {
    auto t = make_triplet(n, n, zH, row1, col1, val, symmetric != 0);
    SpHbMat hb(n, n, true);
    hb.setStructure(t);
    if (val_refresh) {
        auto t2 = make_triplet(n, n, zH, row1, col1, val_refresh, symmetric != 0);
        hb.setMatVal(t2);
    }
    export_csr(hb, rowptr, colidx, out_val, order);
    return hb.EntryNum();
}

// SpHbMat::times (src/SpHbMat.cpp:698-737) on a compressed-row matrix.
void ref_csr_times(int nrow, int ncol, int nnz, const int* rowptr, const int* colidx,
                   const double* val, const double* x, double* y) {
    auto m = std::make_shared<SpHbMat>(nnz, nrow, ncol, true);
    for (int i = 0; i <= nrow; i++) m->setRowIndexAt(i, rowptr[i]);
    for (int i = 0; i < nnz; i++) {
        m->setColIndexAt(i, colidx[i]);
        m->setMatValAt(i, val[i]);
    }
    auto p = std::make_shared<Vector>(ncol, x);
    auto r = std::make_shared<Vector>(nrow);
    if (nnz > 0) m->times(p, r);
    std::memcpy(y, r->values(), sizeof(double) * nrow);
}

}  // extern "C"
