// nlp_eval.cu -- device-side batched NLP evaluation: the straight-line program that
// restartsqp_b200/nl_reader.py derives from a model's expression DAG (f, c, grad f, Jacobian and Lagrangian-Hessian
// triplet values; one thread per instance) is compiled at run time with NVRTC for sm_100a and launched through the
// CUDA runtime's library API.  It replaces, for a batch, what the reference obtains instance by instance from
// Ipopt's AmplTNLP (src/SQPTNLP.cpp:67-132).  NVRTC is loaded with dlopen so that libsqpb200.so itself carries no
// link-time dependency on it.  --fmad=false: the arithmetic is operation for operation the host evaluator's.
#include "../../include/sqpb200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* so = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    bool load(std::string& err) {
        if (so) return true;
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* n : names) { so = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
        if (!so) { err = "cannot load libnvrtc"; return false; }
#define SYM(f) *(void**)(&f) = dlsym(so, "nvrtc" #f); if (!f) { err = "libnvrtc lacks nvrtc" #f; return false; }
        SYM(CreateProgram) SYM(CompileProgram) SYM(GetCUBINSize) SYM(GetCUBIN) SYM(GetProgramLogSize) SYM(GetProgramLog) SYM(DestroyProgram)
#undef SYM
        return true;
    }
};
Nvrtc g_nvrtc;
}  // namespace

struct sqpb200_nlp_s {
    int device = 0, n = 0, m = 0, zJ = 0, zH = 0;
    std::vector<char> cubin;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_all = nullptr, k_fc = nullptr;
    std::string err;
    void* stage = nullptr;  // device staging for host-side callers
    size_t stage_bytes = 0;
    long long launches = 0;
};

static thread_local std::string g_last_nlp_error;
const char* sqpb200_nlp_last_error(void) { return g_last_nlp_error.c_str(); }

int sqpb200_nlp_compile(const char* cuda_source, int n, int m, int zJ, int zH, char* log, int log_len, sqpb200_nlp* out) {
    if (!cuda_source || !out) return SQPB200_ERR_INVALID;
    std::string err;
    if (!g_nvrtc.load(err)) { g_last_nlp_error = err; return SQPB200_ERR_STATE; }
    nvrtcProgram prog = nullptr;
    if (g_nvrtc.CreateProgram(&prog, cuda_source, "nlp_eval.cu", 0, nullptr, nullptr) != 0) { g_last_nlp_error = "nvrtcCreateProgram failed"; return SQPB200_ERR_STATE; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "--std=c++17"};
    int rc = g_nvrtc.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    std::string plog(ls, '\0');
    if (ls > 1) g_nvrtc.GetProgramLog(prog, &plog[0]);
    if (log && log_len > 0) { strncpy(log, plog.c_str(), (size_t)log_len - 1); log[log_len - 1] = 0; }
    if (rc != 0) { g_last_nlp_error = "NVRTC compilation failed: " + plog; g_nvrtc.DestroyProgram(&prog); return SQPB200_ERR_INVALID; }
    sqpb200_nlp h = new sqpb200_nlp_s();
    h->n = n; h->m = m; h->zJ = zJ; h->zH = zH;
    size_t cs = 0;
    g_nvrtc.GetCUBINSize(prog, &cs);
    h->cubin.resize(cs);
    g_nvrtc.GetCUBIN(prog, h->cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    *out = h;
    return 0;
}

long long sqpb200_nlp_cubin_size(sqpb200_nlp h) { return h ? (long long)h->cubin.size() : 0; }

int sqpb200_nlp_load(sqpb200_nlp h, int device) {
    if (!h) return SQPB200_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { g_last_nlp_error = "no CUDA device (no CPU fallback)"; return SQPB200_ERR_CUDA; }
    h->device = device;
    cudaSetDevice(device);
    cudaError_t e = cudaLibraryLoadData(&h->lib, h->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&h->k_all, h->lib, "nlp_eval_all");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&h->k_fc, h->lib, "nlp_eval_fc");
    if (e != cudaSuccess) { g_last_nlp_error = std::string("loading the NLP kernels: ") + cudaGetErrorString(e); return SQPB200_ERR_CUDA; }
    return 0;
}

int sqpb200_nlp_destroy(sqpb200_nlp h) {
    if (!h) return SQPB200_ERR_INVALID;
    if (h->lib) { cudaSetDevice(h->device); cudaLibraryUnload(h->lib); }
    if (h->stage) cudaFree(h->stage);
    delete h;
    return 0;
}

long long sqpb200_nlp_launch_count(sqpb200_nlp h) { return h ? h->launches : 0; }

// which = 0: f, c only; 1: everything.  Host pointers are staged through one device buffer; NULL outputs are skipped.
int sqpb200_nlp_eval(sqpb200_nlp h, int which, int B, const double* x, const double* lam, double* f, double* c, double* grad,
                     double* jac, double* hess, int loc, void* stream_) {
    if (!h || !h->lib || B <= 0 || !x) return SQPB200_ERR_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaSetDevice(h->device);
    const size_t n = h->n, m = h->m, zJ = h->zJ, zH = h->zH, Bz = (size_t)B;
    if (which == 1 && m && !lam) return SQPB200_ERR_INVALID;  // the Lagrangian Hessian needs multipliers (host and device mode)
    const double *dx = x, *dlam = lam;
    double *df = f, *dc = c, *dg = grad, *dj = jac, *dh = hess;
    size_t off[8];
    if (loc == SQPB200_LOC_HOST) {
        size_t o = 0;
        auto take = [&](size_t cnt) { size_t r = o; o += ((cnt * 8 + 255) / 256) * 256; return r; };
        off[0] = take(Bz * n); off[1] = take(Bz * (m ? m : 1)); off[2] = take(Bz); off[3] = take(Bz * (m ? m : 1));
        off[4] = take(Bz * n); off[5] = take(Bz * (zJ ? zJ : 1)); off[6] = take(Bz * (zH ? zH : 1));
        if (o > h->stage_bytes) {
            if (h->stage) { cudaStreamSynchronize(stream); cudaFree(h->stage); }
            if (cudaMalloc(&h->stage, o) != cudaSuccess) { h->stage = nullptr; h->stage_bytes = 0; g_last_nlp_error = "cudaMalloc failed"; return SQPB200_ERR_NOMEM; }
            h->stage_bytes = o;
        }
        char* s = (char*)h->stage;
        cudaMemcpyAsync(s + off[0], x, Bz * n * 8, cudaMemcpyHostToDevice, stream);
        dx = (const double*)(s + off[0]);
        if (which == 1 && m && lam) cudaMemcpyAsync(s + off[1], lam, Bz * m * 8, cudaMemcpyHostToDevice, stream);
        dlam = (const double*)(s + off[1]);
        df = (double*)(s + off[2]); dc = (double*)(s + off[3]); dg = (double*)(s + off[4]); dj = (double*)(s + off[5]); dh = (double*)(s + off[6]);
    } else if (which == 1 && m && !lam) return SQPB200_ERR_INVALID;
    int Bi = B;
    const int block = 128, grid = (B + block - 1) / block;
    cudaError_t e;
    if (which == 0) {
        void* args[] = {&Bi, &dx, &df, &dc};
        e = cudaLaunchKernel((const void*)h->k_fc, dim3(grid), dim3(block), args, 0, stream);
    } else {
        void* args[] = {&Bi, &dx, &dlam, &df, &dc, &dg, &dj, &dh};
        e = cudaLaunchKernel((const void*)h->k_all, dim3(grid), dim3(block), args, 0, stream);
    }
    if (e != cudaSuccess) { g_last_nlp_error = std::string("NLP kernel launch: ") + cudaGetErrorString(e); return SQPB200_ERR_CUDA; }
    h->launches++;
    if (loc == SQPB200_LOC_HOST) {
        if (f) cudaMemcpyAsync(f, df, Bz * 8, cudaMemcpyDeviceToHost, stream);
        if (c && m) cudaMemcpyAsync(c, dc, Bz * m * 8, cudaMemcpyDeviceToHost, stream);
        if (which == 1) {
            if (grad) cudaMemcpyAsync(grad, dg, Bz * n * 8, cudaMemcpyDeviceToHost, stream);
            if (jac && zJ) cudaMemcpyAsync(jac, dj, Bz * zJ * 8, cudaMemcpyDeviceToHost, stream);
            if (hess && zH) cudaMemcpyAsync(hess, dh, Bz * zH * 8, cudaMemcpyDeviceToHost, stream);
        }
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { g_last_nlp_error = std::string("NLP kernel: ") + cudaGetErrorString(e); return SQPB200_ERR_CUDA; }
    }
    return 0;
}
