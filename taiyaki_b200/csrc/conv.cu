// conv.cu -- scatter half of the time-major 1-D convolution backward
// (taiyaki/layers.py:744-850 `Convolution`; forward and the two weight/column
// GEMMs are library GEMMs on a window-gathered matrix, see layers.py
// `_ConvTimeMajor`).  Given the column gradient dcols [T_out][N][C*k] this
// sums, for every input sample, the <= ceil(k/stride) windows that contain it:
//     dx[t][n][c] = sum_{j = (t+pad) mod stride, step stride, j < k}
//                       dcols[(t + pad - j) / stride][n][c*k + j]
// (a gather, so no atomics and a deterministic sum).  HBM-bound: it reads
// dcols once and writes dx once.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ty {

__global__ void __launch_bounds__(256) col2im_tm_kernel(const float *__restrict__ dcols, int Tout,
                                                        int N, int C, int k, int stride, int pad,
                                                        int T, float *__restrict__ dx) {
    const size_t total = (size_t)T * N * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t tn = i / C;
        const int n = (int)(tn % N);
        const int t = (int)(tn / N);
        const int tau = t + pad;
        float acc = 0.f;
        for (int j = tau % stride; j < k; j += stride) {
            const int to = (tau - j) / stride;
            if (to >= 0 && to < Tout && tau - j >= 0)
                acc += dcols[((size_t)to * N + n) * ((size_t)C * k) + (size_t)c * k + j];
        }
        dx[i] = acc;
    }
}

}  // namespace ty

using namespace ty;

extern "C" int ty_col2im_time_major(const float *dcols, int Tout, int N, int C, int k, int stride,
                                    int pad_left, int T, float *dx, void *stream) {
    if (!dcols || !dx || Tout <= 0 || N <= 0 || C <= 0 || k <= 0 || stride <= 0 || T <= 0 ||
        pad_left < 0) {
        set_error("ty_col2im_time_major: bad argument");
        return TY_EINVAL;
    }
    const size_t total = (size_t)T * N * C;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    col2im_tm_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dcols, Tout, N, C, k, stride, pad_left, T, dx);
    return check_launch("col2im_tm_kernel");
}

// ===========================================================================
// Small direct convolutions (stride 1, few channels): the first layers of the
// flip-flop models (taiyaki models/mLstm_flipflop.py: 1 -> 4 and 4 -> 16
// channels, window 5).  As GEMMs these are degenerate -- the weight gradient
// is a [Cout x M] x [M x C*k] product with M = T*N = 256 000 and a 20- or
// 320-element result, 120-150 us each in the library -- so they are written
// directly: one thread per (time, chunk) position, weights in shared memory.
//   forward   z = b + w * x (pre-activation, kept for the backward), a = f(z)
//   dz        dz = da * f'(z)
//   wgrad     dW[co][ci][j] = sum_p dz[p][co] x[p + (j - pad) N][ci],  db = sum_p dz
//   dgrad     dx[p][ci] = sum_{j,co} dz[p - (j - pad) N][co] w[co][ci][j]
// p = t*N + n indexes positions of the time-major tensors, so a shift by one
// time step is a shift by N positions and the chunk index is preserved.
namespace ty {

enum { kActLinear = 0, kActTanh = 1, kActSwish = 2 };

__device__ __forceinline__ float act_fwd(int act, float z) {
    if (act == kActTanh) return tanhf(z);
    if (act == kActSwish) return z / (1.0f + __expf(-z));
    return z;
}
__device__ __forceinline__ float act_bwd(int act, float z) {
    if (act == kActTanh) {
        const float t = tanhf(z);
        return 1.0f - t * t;
    }
    if (act == kActSwish) {
        const float s = 1.0f / (1.0f + __expf(-z));
        return s * (1.0f + z * (1.0f - s));
    }
    return 1.0f;
}

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256) conv_small_fwd_kernel(
    const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
    long long P, int N, int pl, int act, float *__restrict__ z, float *__restrict__ a) {
    __shared__ float ws[COUT * CIN * K];
    __shared__ float bs[COUT];
    for (int i = threadIdx.x; i < COUT * CIN * K; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = b ? b[i] : 0.f;
    __syncthreads();
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
         p += (long long)gridDim.x * blockDim.x) {
        float xin[CIN][K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const long long q = p + (long long)(j - pl) * N;
            const bool ok = q >= 0 && q < P;
#pragma unroll
            for (int ci = 0; ci < CIN; ci++) xin[ci][j] = ok ? x[q * CIN + ci] : 0.f;
        }
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 4) {
            float zz[4], aa[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int co = c0 + u;
                float acc = bs[co];
#pragma unroll
                for (int ci = 0; ci < CIN; ci++)
#pragma unroll
                    for (int j = 0; j < K; j++) acc = fmaf(ws[(co * CIN + ci) * K + j], xin[ci][j], acc);
                zz[u] = acc;
                aa[u] = act_fwd(act, acc);
            }
            *reinterpret_cast<float4 *>(z + p * COUT + c0) = make_float4(zz[0], zz[1], zz[2], zz[3]);
            *reinterpret_cast<float4 *>(a + p * COUT + c0) = make_float4(aa[0], aa[1], aa[2], aa[3]);
        }
    }
}

__global__ void __launch_bounds__(256) conv_dz_kernel(const float4 *__restrict__ da,
                                                      const float4 *__restrict__ z, long long n4,
                                                      int act, float4 *__restrict__ dz) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 g = da[i], v = z[i];
        dz[i] = make_float4(g.x * act_bwd(act, v.x), g.y * act_bwd(act, v.y),
                            g.z * act_bwd(act, v.z), g.w * act_bwd(act, v.w));
    }
}

// Persistent blocks walk tiles of kWgTile positions; the tile's dz rows and the K
// shifted copies of its x rows sit in shared memory and every thread owns one or
// two (co, ci, j) products (plus the COUT bias sums), accumulated in registers over
// all of the block's tiles and added to the result once.
constexpr int kWgTile = 128;

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256) conv_small_wgrad_kernel(
    const float *__restrict__ dz, const float *__restrict__ x, long long P, int N, int pl,
    float *__restrict__ dW, float *__restrict__ db) {
    constexpr int NW = COUT * CIN * K;
    constexpr int NCOMB = NW + COUT;
    constexpr int XROW = kWgTile * CIN + 4;          // +4: taps land in different banks
    constexpr int PER = (NCOMB + 255) / 256;
    __shared__ float dzs[kWgTile * COUT];
    __shared__ float xs[K * XROW];
    float acc[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) acc[u] = 0.f;
    const long long ntile = (P + kWgTile - 1) / kWgTile;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long p0 = tile * kWgTile;
        __syncthreads();
        for (int i = threadIdx.x; i < kWgTile * COUT; i += 256) {
            const long long g = p0 * COUT + i;
            dzs[i] = g < P * COUT ? dz[g] : 0.f;
        }
        for (int i = threadIdx.x; i < K * kWgTile * CIN; i += 256) {
            const int j = i / (kWgTile * CIN), r = i - j * (kWgTile * CIN);
            const long long q = p0 + r / CIN + (long long)(j - pl) * N;
            xs[j * XROW + r] = (q >= 0 && q < P) ? x[q * CIN + r % CIN] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int c = threadIdx.x + u * 256;
            if (c < NW) {
                const int co = c / (CIN * K), rem = c - co * (CIN * K);
                const int ci = rem / K, j = rem - ci * K;
                const float *dp = dzs + co;
                const float *xp = xs + j * XROW + ci;
                float s = 0.f;
#pragma unroll 8
                for (int p = 0; p < kWgTile; p++) s = fmaf(dp[p * COUT], xp[p * CIN], s);
                acc[u] += s;
            } else if (c < NCOMB) {
                const float *dp = dzs + (c - NW);
                float s = 0.f;
#pragma unroll 8
                for (int p = 0; p < kWgTile; p++) s += dp[p * COUT];
                acc[u] += s;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int c = threadIdx.x + u * 256;
        if (c < NW) atomicAdd(dW + c, acc[u]);
        else if (c < NCOMB && db) atomicAdd(db + (c - NW), acc[u]);
    }
}

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256) conv_small_dgrad_kernel(
    const float *__restrict__ dz, const float *__restrict__ w, long long P, int N, int pl,
    float *__restrict__ dx) {
    __shared__ float ws[COUT * CIN * K];
    for (int i = threadIdx.x; i < COUT * CIN * K; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
         p += (long long)gridDim.x * blockDim.x) {
        float acc[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ci++) acc[ci] = 0.f;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const long long q = p - (long long)(j - pl) * N;     // output position that used x[p] at tap j
            if (q >= 0 && q < P) {
#pragma unroll
                for (int c0 = 0; c0 < COUT; c0 += 4) {
                    const float4 d = *reinterpret_cast<const float4 *>(dz + q * COUT + c0);
                    const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int ci = 0; ci < CIN; ci++)
                            acc[ci] = fmaf(dd[u], ws[((c0 + u) * CIN + ci) * K + j], acc[ci]);
                }
            }
        }
#pragma unroll
        for (int ci = 0; ci < CIN; ci++) dx[p * CIN + ci] = acc[ci];
    }
}

// Window gather for the wide convolutions, straight into the GEMM operand type:
// cols[to*N + n][c*k + j] = x[to*stride + j - pad][n][c]  (bf16), column C*k holds
// 1 (the bias rides in the GEMM) and the remaining pad columns 0.  One thread per
// 8 output columns (one 16-byte store).
__global__ void __launch_bounds__(256) im2col_tm_bf16_kernel(
    const float *__restrict__ x, int T, int N, int C, int k, int stride, int pl, int Tout,
    int ld, unsigned short *__restrict__ cols) {
    const int groups = ld / 8;
    const long long total = (long long)Tout * N * groups;
    const int CK = C * k;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long long row = i / groups;
        const int n = (int)(row % N);
        const int to = (int)(row / N);
        unsigned short v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int q = g * 8 + u;
            float f = 0.f;
            if (q < CK) {
                const int c = q / k, j = q - c * k;
                const int t = to * stride + j - pl;
                if (t >= 0 && t < T) f = x[((long long)t * N + n) * C + c];
            } else if (q == CK) {
                f = 1.0f;
            }
            v[u] = __bfloat16_as_ushort(__float2bfloat16(f));
        }
        uint4 o;
        o.x = v[0] | ((unsigned)v[1] << 16); o.y = v[2] | ((unsigned)v[3] << 16);
        o.z = v[4] | ((unsigned)v[5] << 16); o.w = v[6] | ((unsigned)v[7] << 16);
        *reinterpret_cast<uint4 *>(cols + row * ld + g * 8) = o;
    }
}

// col2im with a row stride (the GEMM operand is padded to a multiple of 8 columns)
__global__ void __launch_bounds__(256) col2im_tm_ld_kernel(const float *__restrict__ dcols, int ld,
                                                           int Tout, int N, int C, int k,
                                                           int stride, int pad, int T,
                                                           float *__restrict__ dx) {
    const size_t total = (size_t)T * N * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t tn = i / C;
        const int n = (int)(tn % N);
        const int t = (int)(tn / N);
        const int tau = t + pad;
        float acc = 0.f;
        for (int j = tau % stride; j < k; j += stride) {
            const int to = (tau - j) / stride;
            if (to >= 0 && to < Tout && tau - j >= 0)
                acc += dcols[((size_t)to * N + n) * (size_t)ld + (size_t)c * k + j];
        }
        dx[i] = acc;
    }
}

// ---------------------------------------------------------------------------
// Single-channel strided feature convolution (the first layer of mGru_flipflop.py /
// mGru_cat_mod_flipflop.py: Convolution(1, size, 19, stride=2), layers.py:795 nn.Conv1d) as direct
// fp32 kernels -- as a GEMM it is [T_out*N x 19] . [19 x size]: too narrow for a tensor-core tile,
// and bf16 would round the raw signal.  z[t][n][c] = b[c] + sum_j w[c][j] x[t*stride + j - pad][n].
// Tile = one output time x kIn1Rows chunks: the k x rows signal window is staged in shared memory
// (rows are contiguous in x), thread c keeps its k weights in registers; persistent blocks.
constexpr int kIn1Rows = 32, kIn1MaxK = 32;

template <int K>
__global__ void __launch_bounds__(256) conv_in1_fwd_kernel(const float *__restrict__ x,
                                                           const float *__restrict__ w,
                                                           const float *__restrict__ b, int T, int N,
                                                           int Cout, int stride, int pad_left, int Tout,
                                                           float *__restrict__ z) {
    __shared__ float xs[kIn1MaxK][kIn1Rows];
    const int k = K > 0 ? K : Cout >> 16, cout = K > 0 ? Cout : Cout & 0xffff;
    const int ntile_n = (N + kIn1Rows - 1) / kIn1Rows;
    const int ntiles = Tout * ntile_n;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile / ntile_n, n0 = (tile - t * ntile_n) * kIn1Rows;
        __syncthreads();
        for (int i = threadIdx.x; i < k * kIn1Rows; i += blockDim.x) {
            const int j = i / kIn1Rows, r = i - j * kIn1Rows;
            const int ts = t * stride + j - pad_left, n = n0 + r;
            xs[j][r] = (ts >= 0 && ts < T && n < N) ? x[(size_t)ts * N + n] : 0.f;
        }
        __syncthreads();
        const int rows = min(kIn1Rows, N - n0);
        for (int c = threadIdx.x; c < cout; c += blockDim.x) {
            float wr[K > 0 ? K : kIn1MaxK];
#pragma unroll
            for (int j = 0; j < (K > 0 ? K : kIn1MaxK); j++) wr[j] = j < k ? w[(size_t)c * k + j] : 0.f;
            const float bc = b ? b[c] : 0.f;
            float *zo = z + ((size_t)t * N + n0) * cout + c;
            for (int r = 0; r < rows; r++) {
                float acc = bc;
#pragma unroll
                for (int j = 0; j < (K > 0 ? K : kIn1MaxK); j++)
                    if (K > 0 || j < k) acc = fmaf(wr[j], xs[j][r], acc);     // rows >= k of xs are never written
                zo[(size_t)r * cout] = acc;
            }
        }
    }
}

// dw[c][j] += sum_{t,n} dz[t][n][c] x[t*stride + j - pad][n],  db[c] += sum dz[t][n][c]
template <int K>
__global__ void __launch_bounds__(256) conv_in1_wgrad_kernel(const float *__restrict__ x,
                                                             const float *__restrict__ dz, int T, int N,
                                                             int Cout, int stride, int pad_left, int Tout,
                                                             float *__restrict__ dw, float *__restrict__ db) {
    __shared__ float xs[kIn1MaxK][kIn1Rows];
    const int k = K > 0 ? K : Cout >> 16, cout = K > 0 ? Cout : Cout & 0xffff;
    const int ntile_n = (N + kIn1Rows - 1) / kIn1Rows;
    const int ntiles = Tout * ntile_n;
    // a thread serves channels c = tid, tid + 256, ...: one register set per pass over the tiles
    for (int c0 = 0; c0 < cout; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float acc[K > 0 ? K : kIn1MaxK], accb = 0.f;
#pragma unroll
        for (int j = 0; j < (K > 0 ? K : kIn1MaxK); j++) acc[j] = 0.f;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int t = tile / ntile_n, n0 = (tile - t * ntile_n) * kIn1Rows;
            __syncthreads();
            for (int i = threadIdx.x; i < k * kIn1Rows; i += blockDim.x) {
                const int j = i / kIn1Rows, r = i - j * kIn1Rows;
                const int ts = t * stride + j - pad_left, n = n0 + r;
                xs[j][r] = (ts >= 0 && ts < T && n < N) ? x[(size_t)ts * N + n] : 0.f;
            }
            __syncthreads();
            if (c < cout) {
                const int rows = min(kIn1Rows, N - n0);
                const float *dzo = dz + ((size_t)t * N + n0) * cout + c;
                for (int r = 0; r < rows; r++) {
                    const float d = dzo[(size_t)r * cout];
                    accb += d;
#pragma unroll
                    for (int j = 0; j < (K > 0 ? K : kIn1MaxK); j++)
                        if (K > 0 || j < k) acc[j] = fmaf(d, xs[j][r], acc[j]);
                }
            }
        }
        if (c < cout) {
#pragma unroll
            for (int j = 0; j < (K > 0 ? K : kIn1MaxK); j++)
                if (j < k) atomicAdd(dw + (size_t)c * k + j, acc[j]);
            if (db) atomicAdd(db + c, accb);
        }
    }
}

static inline unsigned grid_for(long long work, int per_block) {
    long long b = (work + per_block - 1) / per_block;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace ty

// which (C, Cout, k) the direct kernels are instantiated for
#define TY_SMALL_CONV_SHAPES(X) X(1, 4, 5) X(4, 16, 5) X(1, 4, 3) X(4, 16, 3) X(1, 8, 5) X(8, 16, 5) X(1, 16, 5) X(2, 8, 5)

extern "C" int ty_conv_small_supported(int C, int Cout, int k) {
#define X(a, b, c) if (C == a && Cout == b && k == c) return 1;
    TY_SMALL_CONV_SHAPES(X)
#undef X
    return 0;
}

extern "C" int ty_conv_small_forward(const float *x, const float *w, const float *b, int T, int N,
                                     int C, int Cout, int k, int pad_left, int act, float *z,
                                     float *a, void *stream) {
    if (!x || !w || !z || !a || T <= 0 || N <= 0 || act < 0 || act > 2) {
        set_error("ty_conv_small_forward: bad argument");
        return TY_EINVAL;
    }
    const long long P = (long long)T * N;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define X(ci, co, kk)                                                                      \
    if (C == ci && Cout == co && k == kk) {                                                \
        conv_small_fwd_kernel<ci, co, kk><<<grid_for(P, 256), 256, 0, s>>>(x, w, b, P, N,   \
                                                                          pad_left, act, z, a); \
        return check_launch("conv_small_fwd_kernel");                                      \
    }
    TY_SMALL_CONV_SHAPES(X)
#undef X
    set_error("ty_conv_small_forward: shape C=%d Cout=%d k=%d not instantiated", C, Cout, k);
    return TY_EINVAL;
}

// dz [T][N][Cout] is scratch the caller provides; dW [Cout][C][k] and db [Cout] are
// ACCUMULATED into (zero them first); dx may be NULL (first layer).
extern "C" int ty_conv_small_backward(const float *da, const float *z, const float *x,
                                      const float *w, int T, int N, int C, int Cout, int k,
                                      int pad_left, int act, float *dz, float *dW, float *db,
                                      float *dx, void *stream) {
    if (!da || !z || !x || !w || !dz || !dW || T <= 0 || N <= 0 || act < 0 || act > 2) {
        set_error("ty_conv_small_backward: bad argument");
        return TY_EINVAL;
    }
    const long long P = (long long)T * N;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define X(ci, co, kk)                                                                          \
    if (C == ci && Cout == co && k == kk) {                                                    \
        const long long n4 = P * co / 4;                                                       \
        conv_dz_kernel<<<grid_for(n4, 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(da), \
                                                         reinterpret_cast<const float4 *>(z),  \
                                                         n4, act,                              \
                                                         reinterpret_cast<float4 *>(dz));      \
        conv_small_wgrad_kernel<ci, co, kk><<<296, 256, 0, s>>>(dz, x, P, N, pad_left, dW, db); \
        if (dx)                                                                                \
            conv_small_dgrad_kernel<ci, co, kk><<<grid_for(P, 256), 256, 0, s>>>(dz, w, P, N,   \
                                                                               pad_left, dx);  \
        return check_launch("conv_small backward kernels");                                    \
    }
    TY_SMALL_CONV_SHAPES(X)
#undef X
    set_error("ty_conv_small_backward: shape C=%d Cout=%d k=%d not instantiated", C, Cout, k);
    return TY_EINVAL;
}

extern "C" int ty_im2col_time_major_bf16(const float *x, int T, int N, int C, int k, int stride,
                                         int pad_left, int Tout, int ld, void *cols_bf16,
                                         void *stream) {
    if (!x || !cols_bf16 || T <= 0 || N <= 0 || C <= 0 || k <= 0 || stride <= 0 || Tout <= 0 ||
        ld % 8 != 0 || ld < C * k + 1) {
        set_error("ty_im2col_time_major_bf16: bad argument (ld must be a multiple of 8, > C*k)");
        return TY_EINVAL;
    }
    const long long total = (long long)Tout * N * (ld / 8);
    im2col_tm_bf16_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, T, N, C, k, stride, pad_left, Tout, ld, static_cast<unsigned short *>(cols_bf16));
    return check_launch("im2col_tm_bf16_kernel");
}

extern "C" int ty_col2im_time_major_ld(const float *dcols, int ld, int Tout, int N, int C, int k,
                                       int stride, int pad_left, int T, float *dx, void *stream) {
    if (!dcols || !dx || Tout <= 0 || N <= 0 || C <= 0 || k <= 0 || stride <= 0 || T <= 0 ||
        pad_left < 0 || ld < C * k) {
        set_error("ty_col2im_time_major_ld: bad argument");
        return TY_EINVAL;
    }
    const size_t total = (size_t)T * N * C;
    col2im_tm_ld_kernel<<<grid_for((long long)total, 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(dcols, ld, Tout, N, C, k, stride,
                                                               pad_left, T, dx);
    return check_launch("col2im_tm_ld_kernel");
}

extern "C" int ty_conv_in1_supported(int C, int Cout, int k) {
    return C == 1 && k >= 1 && k <= kIn1MaxK && Cout >= 1 && Cout <= 0xffff;
}

// grid: persistent blocks, a few per SM
static inline unsigned in1_grid(int Tout, int N) {
    const long long tiles = (long long)Tout * ((N + kIn1Rows - 1) / kIn1Rows);
    return (unsigned)(tiles < 148 * 4 ? (tiles < 1 ? 1 : tiles) : 148 * 4);
}

extern "C" int ty_conv_in1_forward(const float *x, const float *w, const float *b, int T, int N, int Cout,
                                   int k, int stride, int pad_left, int Tout, float *z, void *stream) {
    if (!x || !w || !z || T <= 0 || N <= 0 || stride <= 0 || Tout <= 0 || !ty_conv_in1_supported(1, Cout, k)) {
        set_error("ty_conv_in1_forward: bad argument (T=%d N=%d Cout=%d k=%d stride=%d)", T, N, Cout, k, stride);
        return TY_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (k == 19)
        conv_in1_fwd_kernel<19><<<in1_grid(Tout, N), 256, 0, s>>>(x, w, b, T, N, Cout, stride, pad_left, Tout, z);
    else
        conv_in1_fwd_kernel<0><<<in1_grid(Tout, N), 256, 0, s>>>(x, w, b, T, N, Cout | (k << 16), stride, pad_left,
                                                                Tout, z);
    return check_launch("conv_in1_fwd_kernel");
}

extern "C" int ty_conv_in1_wgrad(const float *x, const float *dz, int T, int N, int Cout, int k, int stride,
                                 int pad_left, int Tout, float *dw, float *db, void *stream) {
    if (!x || !dz || !dw || T <= 0 || N <= 0 || stride <= 0 || Tout <= 0 || !ty_conv_in1_supported(1, Cout, k)) {
        set_error("ty_conv_in1_wgrad: bad argument (T=%d N=%d Cout=%d k=%d stride=%d)", T, N, Cout, k, stride);
        return TY_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (k == 19)
        conv_in1_wgrad_kernel<19><<<in1_grid(Tout, N), 256, 0, s>>>(x, dz, T, N, Cout, stride, pad_left, Tout, dw, db);
    else
        conv_in1_wgrad_kernel<0><<<in1_grid(Tout, N), 256, 0, s>>>(x, dz, T, N, Cout | (k << 16), stride, pad_left,
                                                                  Tout, dw, db);
    return check_launch("conv_in1_wgrad_kernel");
}

