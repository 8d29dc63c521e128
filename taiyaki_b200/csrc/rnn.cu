// rnn.cu -- persistent LSTM / GRU recurrence for sm_100a (forward and BPTT).
// Replaces the cuDNN calls behind taiyaki/layers.py:515 (nn.LSTM in `Lstm`) and
// :633 (nn.GRU in `GruMod`), including the time reversal of `Reverse`
// (layers.py:117-153) which becomes a loop direction instead of two flips.
//
// Decomposition (DESIGN.md "Recurrent kernels"):
//   * The input projection x W_ih^T + b_ih of all time steps is one dense GEMM
//     done by the caller; this file owns the strictly sequential part.
//   * Chunks are independent, so the batch is cut into groups of 8 chunks and
//     each group is given to one thread-block CLUSTER of 8 CTAs (8 SMs).  CTA j
//     of a cluster owns hidden units [j*U, (j+1)*U), U = H/8, for all gates.
//   * The CTA's slice of W_hh (bf16) lives in REGISTERS as mma.sync A fragments
//     for the whole sequence: per step a warp only loads the B operand (h_{t-1},
//     bf16, 8 chunks) from shared memory with ldmatrix, issues its HMMA chain
//     into fp32 accumulators, and applies the gate non-linearities in the
//     accumulator registers (one thread holds all gates of a (unit, chunk)
//     cell; the cell state never leaves registers).
//   * h_t is exchanged inside the cluster through distributed shared memory:
//     st.async (16-byte) into every peer's double-buffered h tile, completion
//     counted by the peer's mbarrier (complete_tx), so a step costs one DSMEM
//     hop and no cluster-wide barrier.
//   * Backward keeps the same ownership.  Each CTA multiplies its slice of the
//     gate gradients by its slice of W_hh^T (again register-resident) and the
//     partial dL/dh_{t-1} tiles are reduce-scattered through DSMEM.
// Arithmetic: bf16 operands, fp32 accumulation for the recurrent product;
// everything else (input projection result, gates, cell state, outputs,
// gradients) is fp32.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ty {

constexpr int kCluster = 8;   // CTAs per cluster (portable maximum)
constexpr int kNB = 8;        // chunks per cluster (= mma N)

enum { kLstm = 0, kGru = 1 };

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
        "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
        : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t raddr, float x, float y, uint32_t rbar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
        "r"(__float_as_uint(x)), "r"(__float_as_uint(y)), "r"(rbar)
        : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 1/(1+2^(-x log2 e)): two SFU ops, relative error ~2^-22
__device__ __forceinline__ float sigmoidf_(float x) {
    return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}
// tanh(x) = 2 sigmoid(2x) - 1: absolute error ~2e-7, saturates cleanly
__device__ __forceinline__ float tanhf_(float x) {
    return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f);
}

struct RnnArgs {
    const float *xproj;    // fwd: [T][N][G*H]
    const float *w_hh;     // [G*H][H]
    int T, N, reverse;
    float *y;              // [T][N][H]
    float *reserve;        // [T][N][NS][H]
    const float *dy;       // bwd: [T][N][H]
    float *dxproj;         // bwd: [T][N][G*H]
    float *dhn;            // bwd GRU: [T][N][H] gradient of the hidden-side n pre-activation
};

template <int CELL> struct Cell;
template <> struct Cell<kLstm> { static constexpr int G = 4, NS = 5; };   // i f g o | c
template <> struct Cell<kGru> { static constexpr int G = 3, NS = 4; };    // r z n | W_hn h

// ---------------------------------------------------------------------------
// Forward.  H multiple of 64, H <= 256.  Threads: 32 * H/64 (warp w owns 8 units).
template <int CELL, int H>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(H / 2, 1)
    rnn_forward_kernel(const RnnArgs a) {
    constexpr int G = Cell<CELL>::G, NS = Cell<CELL>::NS;
    constexpr int U = H / kCluster;     // units per CTA
    constexpr int KT = H / 16;          // k tiles
    constexpr int HS = H + 8;           // padded row (bf16) -> conflict-free ldmatrix
    constexpr int NTHR = H / 2;

    __shared__ __align__(16) __nv_bfloat16 hs[2][kNB][HS];
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / kCluster;
    const int unit = rank * U + warp * 8 + r;      // hidden unit of this thread's cells
    const int T = a.T, N = a.N;
    const int b0 = group * kNB + 2 * q;            // chunk of accumulator column 0
    const bool v0 = b0 < N, v1 = b0 + 1 < N;

    // --- W_hh slice -> A fragments (registers, whole sequence) ---
    uint32_t A[2][KT][4];
#pragma unroll
    for (int m = 0; m < 2; m++) {
        const int glo = 2 * m, ghi = 2 * m + 1;
        const float *wlo = a.w_hh + ((size_t)glo * H + unit) * H;
        const float *whi = a.w_hh + ((size_t)(ghi < G ? ghi : 0) * H + unit) * H;
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
            const int k = 16 * kt + 2 * q;
            A[m][kt][0] = pack_bf16(wlo[k], wlo[k + 1]);
            A[m][kt][2] = pack_bf16(wlo[k + 8], wlo[k + 9]);
            if (ghi < G) {
                A[m][kt][1] = pack_bf16(whi[k], whi[k + 1]);
                A[m][kt][3] = pack_bf16(whi[k + 8], whi[k + 9]);
            } else {
                A[m][kt][1] = 0u;
                A[m][kt][3] = 0u;
            }
        }
    }

    for (int i = tid; i < 2 * kNB * HS; i += NTHR) (&hs[0][0][0])[i] = __float2bfloat16(0.f);
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    cluster_sync_all();

    float cst[2] = {0.f, 0.f};    // LSTM cell state / GRU previous h (fp32)
    uint32_t phase = 0u;          // bit b = parity to wait for on full[b]

    auto tindex = [&](int s) { return a.reverse ? T - 1 - s : s; };
    float xp[G][2], xn[G][2];
    auto load_x = [&](float (&dst)[G][2], int s) {
        if (s < T) {
            const size_t base = ((size_t)tindex(s) * N + b0) * (G * H) + unit;
#pragma unroll
            for (int g = 0; g < G; g++) {
                dst[g][0] = v0 ? __ldg(a.xproj + base + (size_t)g * H) : 0.f;
                dst[g][1] = v1 ? __ldg(a.xproj + base + (size_t)(G * H) + (size_t)g * H) : 0.f;
            }
        }
    };
    load_x(xp, 0);

    const uint32_t hs_base = smem_u32(&hs[0][0][0]);
    // ldmatrix source row for this lane: matrix (lane>>3) -> k offset 8*(lane>>3), row lane&7
    const uint32_t ld_off = (uint32_t)(((lane & 7) * HS + 8 * (lane >> 3)) * 2);

    for (int s = 0; s < T; s++) {
        const int t = tindex(s);
        const int cur = s & 1, nxt = cur ^ 1;
        if (s > 0) {
            mbar_wait(&full[cur], (phase >> cur) & 1u);
            phase ^= 1u << cur;
        }

        float acc[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[m][c][e] = 0.f;

        const uint32_t hcur = hs_base + (uint32_t)(cur * kNB * HS * 2) + ld_off;
#pragma unroll
        for (int kp = 0; kp < KT / 2; kp++) {
            uint32_t bf[4];
            ldmatrix_x4(bf, hcur + kp * 64);
            mma_bf16(acc[0][(2 * kp) & 3], A[0][2 * kp], bf[0], bf[1]);
            mma_bf16(acc[1][(2 * kp) & 3], A[1][2 * kp], bf[0], bf[1]);
            mma_bf16(acc[0][(2 * kp + 1) & 3], A[0][2 * kp + 1], bf[2], bf[3]);
            mma_bf16(acc[1][(2 * kp + 1) & 3], A[1][2 * kp + 1], bf[2], bf[3]);
        }
        float pre[2][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                pre[m][e] = (acc[m][0][e] + acc[m][1][e]) + (acc[m][2][e] + acc[m][3][e]);

        float hnew[2];
#pragma unroll
        for (int col = 0; col < 2; col++) {
            const bool valid = col == 0 ? v0 : v1;
            const size_t cell = (size_t)t * N + b0 + col;
            float *res = a.reserve + (cell * NS) * H + unit;
            if (CELL == kLstm) {
                const float gi = sigmoidf_(pre[0][col] + xp[0][col]);
                const float gf = sigmoidf_(pre[0][2 + col] + xp[1][col]);
                const float gg = tanhf_(pre[1][col] + xp[2][col]);
                const float go = sigmoidf_(pre[1][2 + col] + xp[3][col]);
                const float c = gf * cst[col] + gi * gg;
                cst[col] = c;
                hnew[col] = go * tanhf_(c);
                if (valid) {
                    res[0] = gi; res[H] = gf; res[2 * H] = gg; res[3 * H] = go; res[4 * H] = c;
                }
            } else {
                const float hr = pre[0][col], hz = pre[0][2 + col], hn = pre[1][col];
                const float gr = sigmoidf_(hr + xp[0][col]);
                const float gz = sigmoidf_(hz + xp[1][col]);
                const float gn = tanhf_(xp[2][col] + gr * hn);
                const float h = (1.0f - gz) * gn + gz * cst[col];
                cst[col] = h;
                hnew[col] = h;
                if (valid) {
                    res[0] = gr; res[H] = gz; res[2 * H] = gn; res[3 * H] = hn;
                }
            }
            if (valid) a.y[cell * H + unit] = hnew[col];
        }
        load_x(xn, s + 1);     // after xp was consumed: a full step to land

        if (s + 1 < T) {
            // own slice of h_t into the local copy of the next buffer
            hs[nxt][2 * q][unit] = __float2bfloat16(hnew[0]);
            hs[nxt][2 * q + 1][unit] = __float2bfloat16(hnew[1]);
            __syncthreads();
            if (tid == 0) mbar_arrive_expect_tx(&full[nxt], (kCluster - 1) * kNB * U * 2);
            // ... and into the 7 peers: 16-byte pieces, completion on the peer's mbarrier
            constexpr int PPR = U / 8;                 // pieces per row
            constexpr int PIECES = kNB * PPR;
            const uint32_t bar_local = smem_u32(&full[nxt]);
            for (int idx = tid; idx < (kCluster - 1) * PIECES; idx += NTHR) {
                const int pi = idx / PIECES;
                const uint32_t peer = pi + (pi >= (int)rank);
                const int piece = idx - pi * PIECES;
                const int n = piece / PPR, w8 = piece - n * PPR;
                const __nv_bfloat16 *src = &hs[nxt][n][rank * U + 8 * w8];
                const uint4 v = *reinterpret_cast<const uint4 *>(src);
                st_async_v4(mapa(smem_u32(src), peer), v, mapa(bar_local, peer));
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++) { xp[g][0] = xn[g][0]; xp[g][1] = xn[g][1]; }
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Backward (BPTT).  Same ownership as forward.  K dimension of the per-step
// product is this CTA's G*U gate rows (padded to a multiple of 16); M is all H
// hidden units; the partial results are reduce-scattered to the owners.
template <int CELL, int H>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(H / 2, 1)
    rnn_backward_kernel(const RnnArgs a) {
    constexpr int G = Cell<CELL>::G, NS = Cell<CELL>::NS;
    constexpr int U = H / kCluster;
    constexpr int KL = (G * U + 15) / 16 * 16;   // local gate rows, padded
    constexpr int KT = KL / 16;
    constexpr int DS = KL + 8;                   // padded row (bf16)
    constexpr int NTHR = H / 2;
    constexpr int MT = 4;                        // m tiles (16 units) per warp: (H/16)/(H/64)

    __shared__ __align__(16) __nv_bfloat16 ds[kNB][DS];
    __shared__ __align__(16) float rs[2][kCluster][U][kNB];
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / kCluster;
    const int ul = warp * 8 + r;                  // local unit of this thread's cells
    const int unit = rank * U + ul;
    const int T = a.T, N = a.N;
    const int b0 = group * kNB + 2 * q;
    const bool v0 = b0 < N, v1 = b0 + 1 < N;

    // --- W_hh^T slice -> A fragments: A[m = hidden unit][k = local gate row] ---
    // local gate row kl = g*U + u  <->  W_hh row g*H + rank*U + u
    uint32_t A[MT][KT][4];
    auto wrow = [&](int kl, int h) -> float {
        if (kl >= G * U) return 0.f;
        const int g = kl / U, u = kl - g * U;
        return a.w_hh[((size_t)g * H + rank * U + u) * H + h];
    };
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const int h0 = 64 * warp + 16 * mt + r;
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
            const int k = 16 * kt + 2 * q;
            A[mt][kt][0] = pack_bf16(wrow(k, h0), wrow(k + 1, h0));
            A[mt][kt][1] = pack_bf16(wrow(k, h0 + 8), wrow(k + 1, h0 + 8));
            A[mt][kt][2] = pack_bf16(wrow(k + 8, h0), wrow(k + 9, h0));
            A[mt][kt][3] = pack_bf16(wrow(k + 8, h0 + 8), wrow(k + 9, h0 + 8));
        }
    }
    for (int i = tid; i < kNB * DS; i += NTHR) (&ds[0][0])[i] = __float2bfloat16(0.f);
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    cluster_sync_all();

    uint32_t phase = 0u;
    float carry[2] = {0.f, 0.f};   // LSTM: dL/dc carried back; GRU: z * dL/dh carried back
    auto tindex = [&](int sf) { return a.reverse ? T - 1 - sf : sf; };   // forward step -> time

    // per-step inputs, prefetched one step ahead
    struct In { float sv[NS][2]; float dy[2]; float prev[2]; };
    In cur_in, nxt_in;
    auto load_in = [&](In &d, int s) {
        if (s >= T) return;
        const int sf = T - 1 - s;                 // forward step being differentiated
        const int t = tindex(sf);
#pragma unroll
        for (int col = 0; col < 2; col++) {
            const bool valid = col == 0 ? v0 : v1;
            const size_t cell = (size_t)t * N + b0 + col;
#pragma unroll
            for (int k = 0; k < NS; k++)
                d.sv[k][col] = valid ? __ldg(a.reserve + (cell * NS + k) * H + unit) : 0.f;
            d.dy[col] = valid ? __ldg(a.dy + cell * H + unit) : 0.f;
            float pv = 0.f;
            if (valid && sf > 0) {
                const size_t pcell = (size_t)tindex(sf - 1) * N + b0 + col;
                pv = CELL == kLstm ? __ldg(a.reserve + (pcell * NS + 4) * H + unit)   // c_{t-1}
                                   : __ldg(a.y + pcell * H + unit);                   // h_{t-1}
            }
            d.prev[col] = pv;
        }
    };
    load_in(cur_in, 0);

    const uint32_t ds_base = smem_u32(&ds[0][0]);
    const uint32_t ld_off = (uint32_t)(((lane & 7) * DS + 8 * (lane >> 3)) * 2);

    for (int s = 0; s < T; s++) {
        const int sf = T - 1 - s;
        const int t = tindex(sf);
        const int cur = s & 1, nxt = cur ^ 1;
        float dh[2] = {cur_in.dy[0], cur_in.dy[1]};
        if (s > 0) {
            mbar_wait(&full[cur], (phase >> cur) & 1u);
            phase ^= 1u << cur;
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int j = 0; j < kCluster; j++) {
                const float2 v = *reinterpret_cast<const float2 *>(&rs[cur][j][ul][2 * q]);
                sx += v.x; sy += v.y;
            }
            dh[0] += sx; dh[1] += sy;
        }

        float dg[G][2];
#pragma unroll
        for (int col = 0; col < 2; col++) {
            const bool valid = col == 0 ? v0 : v1;
            const size_t cell = (size_t)t * N + b0 + col;
            if (CELL == kLstm) {
                const float gi = cur_in.sv[0][col], gf = cur_in.sv[1][col], gg = cur_in.sv[2][col],
                            go = cur_in.sv[3][col], c = cur_in.sv[4][col];
                const float tc = tanhf_(c);
                const float d = dh[col];
                const float dc = carry[col] + d * go * (1.0f - tc * tc);
                dg[3][col] = d * tc * go * (1.0f - go);
                dg[0][col] = dc * gg * gi * (1.0f - gi);
                dg[2][col] = dc * gi * (1.0f - gg * gg);
                dg[1][col] = dc * cur_in.prev[col] * gf * (1.0f - gf);
                carry[col] = dc * gf;
            } else {
                const float gr = cur_in.sv[0][col], gz = cur_in.sv[1][col], gn = cur_in.sv[2][col],
                            hn = cur_in.sv[3][col];
                const float d = dh[col] + carry[col];
                const float dn = d * (1.0f - gz) * (1.0f - gn * gn);      // d n_pre
                dg[1][col] = d * (cur_in.prev[col] - gn) * gz * (1.0f - gz);
                dg[0][col] = dn * hn * gr * (1.0f - gr);
                dg[2][col] = dn;                                          // x-side n gradient
                carry[col] = d * gz;
                const float dhn = dn * gr;                                // hidden-side n gradient
                if (valid) a.dhn[cell * H + unit] = dhn;
                // the recurrent product uses the hidden-side gradient for gate n
                ds[2 * q + col][2 * U + ul] = __float2bfloat16(dhn);
            }
            if (valid) {
#pragma unroll
                for (int g = 0; g < G; g++)
                    a.dxproj[cell * (G * H) + (size_t)g * H + unit] = dg[g][col];
            }
#pragma unroll
            for (int g = 0; g < G; g++)
                if (!(CELL == kGru && g == 2))
                    ds[2 * q + col][g * U + ul] = __float2bfloat16(dg[g][col]);
        }

        load_in(nxt_in, s + 1);   // after this step's inputs were consumed
        if (s + 1 < T) {
            __syncthreads();
            if (tid == 0) mbar_arrive_expect_tx(&full[nxt], kCluster * U * kNB * 4);
            float acc[MT][2][4];
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[mt][c][e] = 0.f;
#pragma unroll
            for (int kp = 0; kp < KT / 2; kp++) {
                uint32_t bf[4];
                ldmatrix_x4(bf, ds_base + ld_off + kp * 64);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    mma_bf16(acc[mt][0], A[mt][2 * kp], bf[0], bf[1]);
                    mma_bf16(acc[mt][1], A[mt][2 * kp + 1], bf[2], bf[3]);
                }
            }
            if (KT & 1) {   // odd number of k tiles (GRU with small U): last tile alone
                uint32_t bf[4];
                ldmatrix_x4(bf, ds_base + ld_off + (KT / 2) * 64);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) mma_bf16(acc[mt][0], A[mt][KT - 1], bf[0], bf[1]);
            }
            // reduce-scatter: rows of this warp's tiles belong to the CTA owning those units
            const uint32_t rs_nxt = smem_u32(&rs[nxt][rank][0][0]);
            const uint32_t bar_local = smem_u32(&full[nxt]);
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int h = 64 * warp + 16 * mt + 8 * half + r;
                    const uint32_t dest = h / U;
                    const int hl = h - dest * U;
                    const float x = acc[mt][0][2 * half] + acc[mt][1][2 * half];
                    const float y = acc[mt][0][2 * half + 1] + acc[mt][1][2 * half + 1];
                    const uint32_t local = rs_nxt + (uint32_t)((hl * kNB + 2 * q) * 4);
                    st_async_v2(mapa(local, dest), x, y, mapa(bar_local, dest));
                }
            }
        }
        cur_in = nxt_in;
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
template <int CELL>
static int launch_rnn(bool backward, const RnnArgs &a, int H, cudaStream_t s) {
    const int groups = (a.N + kNB - 1) / kNB;
    const dim3 grid(groups * kCluster);
#define TY_RNN(HH)                                                                  \
    case HH:                                                                        \
        if (backward) rnn_backward_kernel<CELL, HH><<<grid, HH / 2, 0, s>>>(a);     \
        else rnn_forward_kernel<CELL, HH><<<grid, HH / 2, 0, s>>>(a);               \
        break;
    switch (H) {
        TY_RNN(64) TY_RNN(128) TY_RNN(192) TY_RNN(256)
        default:
            set_error("ty_rnn: hidden size %d unsupported (need a multiple of 64, <= 256)", H);
            return TY_EINVAL;
    }
#undef TY_RNN
    return check_launch(backward ? "rnn_backward_kernel" : "rnn_forward_kernel");
}

static int check_shape(int T, int N, int H, const void *p0, const void *p1, const void *p2) {
    if (!p0 || !p1 || !p2 || T <= 0 || N <= 0 || H <= 0) {
        set_error("ty_rnn: bad argument (T=%d N=%d H=%d)", T, N, H);
        return TY_EINVAL;
    }
    return TY_OK;
}

}  // namespace ty

using namespace ty;

extern "C" size_t ty_rnn_reserve_bytes(int cell, int T, int N, int H) {
    const int ns = cell == kLstm ? Cell<kLstm>::NS : Cell<kGru>::NS;
    return (size_t)T * N * ns * H * sizeof(float);
}

extern "C" int ty_lstm_forward(const float *xproj, const float *w_hh, int T, int N, int H,
                               int reverse, float *y, void *reserve, void *stream) {
    if (int rc = check_shape(T, N, H, xproj, w_hh, y)) return rc;
    if (!reserve) { set_error("ty_lstm_forward: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.xproj = xproj; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse; a.y = y;
    a.reserve = static_cast<float *>(reserve);
    return launch_rnn<kLstm>(false, a, H, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_lstm_backward(const float *dy, const float *w_hh, int T, int N, int H,
                                int reverse, const float *y, const void *reserve, float *dxproj,
                                void *stream) {
    if (int rc = check_shape(T, N, H, dy, w_hh, dxproj)) return rc;
    if (!reserve) { set_error("ty_lstm_backward: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.reserve = const_cast<float *>(static_cast<const float *>(reserve));
    a.dxproj = dxproj;
    return launch_rnn<kLstm>(true, a, H, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_gru_forward(const float *xproj, const float *w_hh, int T, int N, int H,
                              int reverse, float *y, void *reserve, void *stream) {
    if (int rc = check_shape(T, N, H, xproj, w_hh, y)) return rc;
    if (!reserve) { set_error("ty_gru_forward: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.xproj = xproj; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse; a.y = y;
    a.reserve = static_cast<float *>(reserve);
    return launch_rnn<kGru>(false, a, H, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_gru_backward(const float *dy, const float *w_hh, int T, int N, int H,
                               int reverse, const float *y, const void *reserve, float *dxproj,
                               float *dhn, void *stream) {
    if (int rc = check_shape(T, N, H, dy, w_hh, dxproj)) return rc;
    if (!reserve || !y || !dhn) { set_error("ty_gru_backward: null pointer"); return TY_EINVAL; }
    RnnArgs a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.reserve = const_cast<float *>(static_cast<const float *>(reserve));
    a.dxproj = dxproj; a.dhn = dhn;
    return launch_rnn<kGru>(true, a, H, static_cast<cudaStream_t>(stream));
}
