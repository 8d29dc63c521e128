// rnn_fp32.cu -- fp32 recurrence for ANY hidden size, behind the gate-major entry points
// (ty_lstm_* / ty_gru_* / ty_rnn_*_ex).  Two jobs:
//   * the parity mode of the recurrent layers: every product in fp32 FFMA (no bf16
//     rounding of W_hh or h), accurate expf / tanhf -- what torch.nn.LSTM / nn.GRU compute
//     on the CPU or with TF32 off (taiyaki/layers.py:515,633 -> cuDNN), to ~1e-6 per step;
//     the tests pin the fast bf16 tensor-core kernels (rnn_ws.cu) against THIS and this
//     against torch at T = 800;
//   * the recurrence for hidden sizes the cluster kernels do not cover (they need a
//     multiple of 64 up to 256): the reference trains size 96 "fast" models and defaults to
//     384 (bin/_bin_argparse.py:16, README.md:354-359).
// Not a speed path: a CTA owns up to 8 chunks and streams the whole W_hh from L2 every
// step (G H^2 x 4 bytes), i.e. ~5 ms per layer pass at H = 256, T = 800 -- the ballpark of
// the cuDNN call it replaces, 10x the cluster kernels.
//
// Layouts: xproj / dxproj [T][N][G*H] gate-major fp32 (PyTorch gate order i f g o / r z n),
// y [T][N][H]; reserve = gates [T][N][4][H] (LSTM i f g o after the non-linearities; GRU
// r z n and W_hn h) followed by the cell state [T][N][H] (LSTM).
#include <cuda_bf16.h>

#include "common.cuh"

namespace ty {

constexpr int kFNB = 8;          // chunks per CTA
constexpr int kFThreads = 512;

struct RnnF32Args {
    const float *xproj, *bias, *w_hh;
    int T, N, H, reverse;
    float *y;
    __nv_bfloat16 *y16;
    float *gates, *cstate;
    const float *dy;
    float *dx;                   // [T][N][G*H] fp32 (or null)
    __nv_bfloat16 *dx16;         // same as bf16 (or null)
    float *dhn;                  // GRU: [T][N][H]
    __nv_bfloat16 *dhn16;
    float *dbias;
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// pre[n][r] = sum_k W[r][k] h[n][k] for the CTA's chunks; rows striped over the threads
template <int G>
__device__ __forceinline__ void matvec_rows(const float *__restrict__ W, const float *hs, float *pre, int H,
                                            int tid) {
    const int rows = G * H;
    for (int r = tid; r < rows; r += kFThreads) {
        float acc[kFNB];
#pragma unroll
        for (int n = 0; n < kFNB; n++) acc[n] = 0.f;
        const float *w = W + (size_t)r * H;
        for (int k = 0; k < H; k += 4) {
            const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + k));
#pragma unroll
            for (int n = 0; n < kFNB; n++) {
                const float4 hv = *reinterpret_cast<const float4 *>(hs + n * H + k);
                acc[n] = fmaf(wv.x, hv.x, acc[n]);
                acc[n] = fmaf(wv.y, hv.y, acc[n]);
                acc[n] = fmaf(wv.z, hv.z, acc[n]);
                acc[n] = fmaf(wv.w, hv.w, acc[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < kFNB; n++) pre[n * rows + r] = acc[n];
    }
}

// out[n][j] = sum_r W[r][j] d[n][r]: columns striped over the threads (coalesced in j)
template <int G>
__device__ __forceinline__ void matvec_cols(const float *__restrict__ W, const float *ds, float *out, int H,
                                            int tid) {
    const int rows = G * H;
    for (int j = tid; j < H; j += kFThreads) {
        float acc[kFNB];
#pragma unroll
        for (int n = 0; n < kFNB; n++) acc[n] = 0.f;
        for (int r = 0; r < rows; r++) {
            const float wv = __ldg(W + (size_t)r * H + j);
#pragma unroll
            for (int n = 0; n < kFNB; n++) acc[n] = fmaf(wv, ds[n * rows + r], acc[n]);
        }
#pragma unroll
        for (int n = 0; n < kFNB; n++) out[n * H + j] = acc[n];
    }
}

// shared memory: hs [8][H] | pre [8][G*H] (forward)      dh [8][H] | ds [8][G*H] | dc [8][H] (backward)
template <int CELL>
__global__ void __launch_bounds__(kFThreads, 1) rnn_fp32_forward_kernel(const RnnF32Args a) {
    constexpr int G = CELL == 0 ? 4 : 3;
    extern __shared__ __align__(16) float sm[];
    const int H = a.H, T = a.T, N = a.N, tid = threadIdx.x;
    float *hs = sm, *pre = sm + kFNB * H;
    const int n0 = blockIdx.x * kFNB;
    const int nv = min(kFNB, N - n0);
    for (int i = tid; i < kFNB * H; i += kFThreads) hs[i] = 0.f;
    // cell state of this thread's cells stays in shared memory too (any H)
    float *cs = pre + kFNB * G * H;
    for (int i = tid; i < kFNB * H; i += kFThreads) cs[i] = 0.f;
    __syncthreads();
    for (int s = 0; s < T; s++) {
        const int t = a.reverse ? T - 1 - s : s;
        matvec_rows<G>(a.w_hh, hs, pre, H, tid);
        __syncthreads();
        for (int i = tid; i < nv * H; i += kFThreads) {
            const int n = i / H, u = i - n * H;
            const size_t cell = ((size_t)t * N + n0 + n) * H + u;
            const float *xp = a.xproj + ((size_t)t * N + n0 + n) * G * H;
            const float *pr = pre + n * G * H;
            float b[4] = {0.f, 0.f, 0.f, 0.f};
            if (a.bias) {
#pragma unroll
                for (int g = 0; g < G; g++) b[g] = a.bias[g * H + u];
            }
            float h;
            float *gt = a.gates + ((size_t)t * N + n0 + n) * 4 * H;
            if (CELL == 0) {
                const float gi = sigm(xp[u] + b[0] + pr[u]);
                const float gf = sigm(xp[H + u] + b[1] + pr[H + u]);
                const float gg = tanhf(xp[2 * H + u] + b[2] + pr[2 * H + u]);
                const float go = sigm(xp[3 * H + u] + b[3] + pr[3 * H + u]);
                const float c = gf * cs[i] + gi * gg;
                cs[i] = c;
                h = go * tanhf(c);
                gt[u] = gi; gt[H + u] = gf; gt[2 * H + u] = gg; gt[3 * H + u] = go;
                a.cstate[cell] = c;
            } else {
                const float hn = pr[2 * H + u];
                const float gr = sigm(xp[u] + b[0] + pr[u]);
                const float gz = sigm(xp[H + u] + b[1] + pr[H + u]);
                const float gn = tanhf(xp[2 * H + u] + b[2] + gr * hn);
                h = (1.0f - gz) * gn + gz * hs[i];
                gt[u] = gr; gt[H + u] = gz; gt[2 * H + u] = gn; gt[3 * H + u] = hn;
            }
            a.y[cell] = h;
            if (a.y16) a.y16[cell] = __float2bfloat16(h);
            cs[kFNB * H + i] = h;           // staged: hs is still being read by nobody, but keep the
                                            // update after the barrier for clarity
        }
        __syncthreads();
        for (int i = tid; i < nv * H; i += kFThreads) hs[i] = cs[kFNB * H + i];
        __syncthreads();
    }
}

template <int CELL>
__global__ void __launch_bounds__(kFThreads, 1) rnn_fp32_backward_kernel(const RnnF32Args a) {
    constexpr int G = CELL == 0 ? 4 : 3;
    extern __shared__ __align__(16) float sm[];
    const int H = a.H, T = a.T, N = a.N, tid = threadIdx.x;
    float *dh = sm, *ds = sm + kFNB * H, *dc = ds + kFNB * G * H;
    const int n0 = blockIdx.x * kFNB;
    const int nv = min(kFNB, N - n0);
    for (int i = tid; i < kFNB * H; i += kFThreads) { dh[i] = 0.f; dc[i] = 0.f; }
    for (int i = tid; i < kFNB * G * H; i += kFThreads) ds[i] = 0.f;
    __syncthreads();
    for (int s = T - 1; s >= 0; s--) {
        const int t = a.reverse ? T - 1 - s : s;
        const int tp = a.reverse ? t + 1 : t - 1;          // time index of the previous step
        for (int i = tid; i < nv * H; i += kFThreads) {
            const int n = i / H, u = i - n * H;
            const size_t row = (size_t)t * N + n0 + n;
            const size_t cell = row * H + u;
            const float *gt = a.gates + row * 4 * H;
            const float dht = a.dy[cell] + dh[i];
            float d[4] = {0.f, 0.f, 0.f, 0.f};            // gradient of the x-side pre-activations
            float *dsn = ds + n * G * H;
            if (CELL == 0) {
                const float gi = gt[u], gf = gt[H + u], gg = gt[2 * H + u], go = gt[3 * H + u];
                const float c = a.cstate[cell];
                const float cp = s > 0 ? a.cstate[((size_t)tp * N + n0 + n) * H + u] : 0.f;
                const float tc = tanhf(c);
                const float dct = dc[i] + dht * go * (1.0f - tc * tc);
                d[0] = dct * gg * gi * (1.0f - gi);
                d[1] = dct * cp * gf * (1.0f - gf);
                d[2] = dct * gi * (1.0f - gg * gg);
                d[3] = dht * tc * go * (1.0f - go);
                dc[i] = dct * gf;
                dh[i] = 0.f;
#pragma unroll
                for (int g = 0; g < 4; g++) dsn[g * H + u] = d[g];
            } else {
                const float gr = gt[u], gz = gt[H + u], gn = gt[2 * H + u], hn = gt[3 * H + u];
                const float hp = s > 0 ? a.y[((size_t)tp * N + n0 + n) * H + u] : 0.f;
                const float dn = dht * (1.0f - gz) * (1.0f - gn * gn);
                d[0] = dn * hn * gr * (1.0f - gr);
                d[1] = dht * (hp - gn) * gz * (1.0f - gz);
                d[2] = dn;
                const float dhid_n = dn * gr;              // gradient of W_hn h
                dh[i] = dht * gz;                          // direct path h_{t-1} -> h_t
                dsn[u] = d[0]; dsn[H + u] = d[1]; dsn[2 * H + u] = dhid_n;
                if (a.dhn) a.dhn[cell] = dhid_n;
                if (a.dhn16) a.dhn16[cell] = __float2bfloat16(dhid_n);
            }
            float *dxr = a.dx ? a.dx + row * G * H : nullptr;
            __nv_bfloat16 *dxr16 = a.dx16 ? a.dx16 + row * G * H : nullptr;
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (dxr) dxr[g * H + u] = d[g];
                if (dxr16) dxr16[g * H + u] = __float2bfloat16(d[g]);
            }
            if (a.dbias) {
#pragma unroll
                for (int g = 0; g < G; g++) atomicAdd(a.dbias + g * H + u, d[g]);
            }
        }
        __syncthreads();
        // dL/dh_{t-1} += W_hh^T dgates (hidden-side gradients)
        if (s > 0) {
            float *tmp = dc + kFNB * H;
            matvec_cols<G>(a.w_hh, ds, tmp, H, tid);
            __syncthreads();
            for (int i = tid; i < nv * H; i += kFThreads) dh[i] += tmp[i];
        }
        __syncthreads();
    }
}

static int launch_fp32(int cell, bool backward, const RnnF32Args &a, cudaStream_t s) {
    const int G = cell == 0 ? 4 : 3;
    const size_t smem = backward ? (size_t)kFNB * a.H * (3 + G) * sizeof(float)
                                 : (size_t)kFNB * a.H * (3 + G) * sizeof(float);
    if (a.H % 4 != 0 || smem > 220 * 1024) {
        set_error("ty_rnn (fp32 path): hidden size %d unsupported (multiple of 4, at most %d)", a.H,
                  (int)(220 * 1024 / (kFNB * (3 + G) * sizeof(float))) / 4 * 4);
        return TY_EINVAL;
    }
    const int grid = (a.N + kFNB - 1) / kFNB;
#define TY_F32(K)                                                                              \
    do {                                                                                       \
        cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        K<<<grid, kFThreads, smem, s>>>(a);                                                    \
    } while (0)
    if (!backward) {
        if (cell == 0) TY_F32(rnn_fp32_forward_kernel<0>); else TY_F32(rnn_fp32_forward_kernel<1>);
    } else {
        if (cell == 0) TY_F32(rnn_fp32_backward_kernel<0>); else TY_F32(rnn_fp32_backward_kernel<1>);
    }
#undef TY_F32
    return check_launch(backward ? "rnn_fp32_backward_kernel" : "rnn_fp32_forward_kernel");
}

static int check_shape(int T, int N, int H, const void *a, const void *b, const void *c) {
    if (!a || !b || !c || T <= 0 || N <= 0 || H <= 0) {
        set_error("ty_rnn: bad argument (T=%d N=%d H=%d)", T, N, H);
        return TY_EINVAL;
    }
    return TY_OK;
}

}  // namespace ty

using namespace ty;

extern "C" size_t ty_rnn_reserve_bytes(int cell, int T, int N, int H) {
    (void)cell;
    return (size_t)T * N * 5 * H * sizeof(float);
}

extern "C" int ty_rnn_forward_ex(int cell, const float *xproj, const float *bias, const float *w_hh, int T,
                                 int N, int H, int reverse, float *y, void *y_bf16, void *reserve,
                                 void *stream) {
    if (int rc = check_shape(T, N, H, xproj, w_hh, y)) return rc;
    if (!reserve || (cell != 0 && cell != 1)) { set_error("ty_rnn_forward_ex: bad argument"); return TY_EINVAL; }
    RnnF32Args a{};
    a.xproj = xproj; a.bias = bias; a.w_hh = w_hh; a.T = T; a.N = N; a.H = H; a.reverse = reverse;
    a.y = y; a.y16 = static_cast<__nv_bfloat16 *>(y_bf16);
    a.gates = static_cast<float *>(reserve);
    a.cstate = a.gates + (size_t)T * N * 4 * H;
    return launch_fp32(cell, false, a, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_rnn_backward_ex(int cell, const float *dy, const float *w_hh, int T, int N, int H,
                                  int reverse, const float *y, const void *reserve, void *dxproj, void *dhn,
                                  int grads_bf16, float *dbias, void *stream) {
    if (int rc = check_shape(T, N, H, dy, w_hh, dxproj)) return rc;
    if (!reserve || (cell != 0 && cell != 1) || (cell == 1 && (!y || !dhn))) {
        set_error("ty_rnn_backward_ex: bad argument");
        return TY_EINVAL;
    }
    RnnF32Args a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.H = H; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.gates = const_cast<float *>(static_cast<const float *>(reserve));
    a.cstate = a.gates + (size_t)T * N * 4 * H;
    if (grads_bf16) {
        a.dx16 = static_cast<__nv_bfloat16 *>(dxproj);
        a.dhn16 = static_cast<__nv_bfloat16 *>(dhn);
    } else {
        a.dx = static_cast<float *>(dxproj);
        a.dhn = static_cast<float *>(dhn);
    }
    a.dbias = dbias;
    return launch_fp32(cell, true, a, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_lstm_forward(const float *xproj, const float *bias, const float *w_hh, int T, int N, int H,
                               int reverse, float *y, void *reserve, void *stream) {
    return ty_rnn_forward_ex(0, xproj, bias, w_hh, T, N, H, reverse, y, nullptr, reserve, stream);
}
extern "C" int ty_lstm_backward(const float *dy, const float *w_hh, int T, int N, int H, int reverse,
                                const float *y, const void *reserve, float *dxproj, float *dbias,
                                void *stream) {
    return ty_rnn_backward_ex(0, dy, w_hh, T, N, H, reverse, y, reserve, dxproj, nullptr, 0, dbias, stream);
}
extern "C" int ty_gru_forward(const float *xproj, const float *bias, const float *w_hh, int T, int N, int H,
                              int reverse, float *y, void *reserve, void *stream) {
    return ty_rnn_forward_ex(1, xproj, bias, w_hh, T, N, H, reverse, y, nullptr, reserve, stream);
}
extern "C" int ty_gru_backward(const float *dy, const float *w_hh, int T, int N, int H, int reverse,
                               const float *y, const void *reserve, float *dxproj, float *dhn, float *dbias,
                               void *stream) {
    return ty_rnn_backward_ex(1, dy, w_hh, T, N, H, reverse, y, reserve, dxproj, dhn, 0, dbias, stream);
}
