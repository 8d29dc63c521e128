// capi.cu -- error plumbing and the HOST-POINTER drop-ins that carry the
// reference's exact C ABI (taiyaki/ctc/libctc.pxd:3-25).  A maintainer can
// link taiyaki/ctc/ctc.pyx against libtaiyaki_b200.so instead of compiling
// c_crf_flipflop.c / c_cat_mod_flipflop.c; see INTEGRATION.md.
//
// These wrappers own a grow-only device pool and a private stream; they copy
// inputs host->device, run the same kernels as the ty_* device entry points
// and copy score / grad back, synchronously, like the reference's functions.
// There is no CPU fallback: if CUDA fails the outputs are filled with NaN
// (the reference's own out-of-memory convention, c_crf_flipflop.c:278-282)
// and the reason is kept in ty_last_error_string().
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace ty {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return TY_ECUDA;
    }
    return TY_OK;
}

// grow-only device pool for the host-pointer ABI
struct HostPool {
    std::mutex mu;
    void *buf = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr;
    bool reserve(size_t n) {
        if (!stream && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess)
            return false;
        if (n <= cap) return true;
        if (buf) cudaFree(buf);
        buf = nullptr; cap = 0;
        if (cudaMalloc(&buf, n) != cudaSuccess) return false;
        cap = n;
        return true;
    }
};
static HostPool g_pool;

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static void fill_nan(float *p, size_t n) {
    for (size_t i = 0; i < n; i++) p[i] = NAN;
}

static void host_crf(const float *logprob, size_t ntrans, size_t nblk, size_t nbatch,
                     const size_t *moveidxs, const size_t *stayidxs, const size_t *modmoveidxs,
                     const float *modmovefacts, const int32_t *seqlen, float *score,
                     float *grad) {
    size_t total = 0;
    int max_len = 0;
    for (size_t b = 0; b < nbatch; b++) {
        if (seqlen[b] < 0) {
            set_error("host ABI: negative seqlen");
            fill_nan(score, nbatch);
            return;
        }
        total += (size_t)seqlen[b];
        if (seqlen[b] > max_len) max_len = seqlen[b];
    }
    // The reference's Python packs one move entry per non-final position
    // (ctc.pyx:127-129); the C side reads chunk b at sum(seqlen[:b]) - b
    // (c_crf_flipflop.c:479).  Copy exactly what the caller owns.
    size_t nonempty = 0;
    for (size_t b = 0; b < nbatch; b++) nonempty += seqlen[b] > 0;
    const size_t nmove = total - nonempty;
    const bool mod = modmoveidxs != nullptr;
    const size_t nelem = nblk * nbatch * ntrans;
    const size_t ws_bytes =
        ty_crf_flipflop_workspace_bytes((int)ntrans, (int)nblk, (int)nbatch, max_len, grad != nullptr);

    std::lock_guard<std::mutex> lock(g_pool.mu);
    size_t o = 0;
    const size_t o_lp = o; o += up256(nelem * sizeof(float));
    const size_t o_gr = o; o += grad ? up256(nelem * sizeof(float)) : 0;
    const size_t o_sc = o; o += up256(nbatch * sizeof(float));
    const size_t o_sl = o; o += up256(nbatch * sizeof(int32_t));
    const size_t o_st = o; o += up256((total + 1) * sizeof(int32_t));
    const size_t o_mv = o; o += up256((total + 1) * sizeof(int32_t));
    const size_t o_mm = o; o += mod ? up256((total + 1) * sizeof(int32_t)) : 0;
    const size_t o_mf = o; o += mod ? up256((total + 1) * sizeof(float)) : 0;
    const size_t o_ws = o; o += up256(ws_bytes);
    bool ok = g_pool.reserve(o);
    if (ok) {
        char *d = static_cast<char *>(g_pool.buf);
        cudaStream_t s = g_pool.stream;
        std::vector<int32_t> st(total + 1), mv(total + 1), mm;
        for (size_t i = 0; i < total; i++) st[i] = (int32_t)stayidxs[i];
        for (size_t i = 0; i < nmove; i++) mv[i] = (int32_t)moveidxs[i];
        if (mod) {
            mm.resize(total + 1);
            for (size_t i = 0; i < nmove; i++) mm[i] = (int32_t)modmoveidxs[i];
        }
        cudaMemcpyAsync(d + o_lp, logprob, nelem * sizeof(float), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d + o_sl, seqlen, nbatch * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d + o_st, st.data(), (total + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d + o_mv, mv.data(), (total + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        if (mod) {
            cudaMemcpyAsync(d + o_mm, mm.data(), (total + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s);
            cudaMemcpyAsync(d + o_mf, modmovefacts, nmove * sizeof(float), cudaMemcpyHostToDevice, s);
        }
        // staging vectors must outlive the async copies from pageable memory
        cudaStreamSynchronize(s);
        int rc = ty_crf_flipflop(
            reinterpret_cast<float *>(d + o_lp), (int)ntrans, (int)nblk, (int)nbatch,
            reinterpret_cast<int32_t *>(d + o_mv), reinterpret_cast<int32_t *>(d + o_st),
            mod ? reinterpret_cast<int32_t *>(d + o_mm) : nullptr,
            mod ? reinterpret_cast<float *>(d + o_mf) : nullptr,
            reinterpret_cast<int32_t *>(d + o_sl), max_len, 1.0f, (int)ntrans, 1.0f,
            reinterpret_cast<float *>(d + o_sc), 1.0f,
            grad ? reinterpret_cast<float *>(d + o_gr) : nullptr, d + o_ws, ws_bytes, s);
        ok = rc == TY_OK;
        if (ok) {
            cudaMemcpyAsync(score, d + o_sc, nbatch * sizeof(float), cudaMemcpyDeviceToHost, s);
            if (grad)
                cudaMemcpyAsync(grad, d + o_gr, nelem * sizeof(float), cudaMemcpyDeviceToHost, s);
            cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) {
                set_error("host ABI: %s", cudaGetErrorString(e));
                ok = false;
            }
        }
    } else {
        set_error("host ABI: cannot reserve %zu device bytes: %s", o,
                  cudaGetErrorString(cudaGetLastError()));
    }
    if (!ok) {
        fill_nan(score, nbatch);
        if (grad) fill_nan(grad, nelem);
    }
}

}  // namespace ty

using namespace ty;

extern "C" const char *ty_last_error_string(void) { return g_err; }
extern "C" const char *ty_version(void) { return "taiyaki_b200 0.2 (sm_100a)"; }

extern "C" void crf_flipflop_grad(const float *logprob, size_t ntrans, size_t nblk,
                                  size_t nbatch, const size_t *moveidxs,
                                  const size_t *stayidxs, const int32_t *seqlen, float *score,
                                  float *grad) {
    host_crf(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, nullptr, nullptr, seqlen, score,
             grad);
}

extern "C" void crf_flipflop_cost(const float *logprob, size_t ntrans, size_t nblk,
                                  size_t nbatch, const size_t *moveidxs,
                                  const size_t *stayidxs, const int32_t *seqlen, float *score) {
    host_crf(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, nullptr, nullptr, seqlen, score,
             nullptr);
}

extern "C" void cat_mod_flipflop_grad(const float *logprob, size_t ntrans, size_t nblk,
                                      size_t nbatch, const size_t *moveidxs,
                                      const size_t *stayidxs, const size_t *modmoveidxs,
                                      const float *modmovefacts, const int32_t *seqlen,
                                      float *score, float *grad) {
    host_crf(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, modmoveidxs, modmovefacts,
             seqlen, score, grad);
}

extern "C" void cat_mod_flipflop_cost(const float *logprob, size_t ntrans, size_t nblk,
                                      size_t nbatch, const size_t *moveidxs,
                                      const size_t *stayidxs, const size_t *modmoveidxs,
                                      const float *modmovefacts, const int32_t *seqlen,
                                      float *score) {
    host_crf(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, modmoveidxs, modmovefacts,
             seqlen, score, nullptr);
}
