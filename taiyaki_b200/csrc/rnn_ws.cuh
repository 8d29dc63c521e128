// rnn_ws.cuh -- warp-specialised persistent LSTM / GRU recurrence for sm_100a
// (forward and BPTT) behind the UNIT-MAJOR entry points ty_rnn_forward_um /
// ty_rnn_backward_um.  They replace the cuDNN calls behind
// taiyaki/layers.py:515 nn.LSTM and :633 nn.GRU, plus the time reversal of
// layers.py:117-153 as a loop direction), same cluster decomposition:
//   * groups of 8 chunks, one cluster of 8 CTAs per group, CTA j owns hidden
//     units [j*U, (j+1)*U), U = H/8, for all gates; its W_hh slice lives in
//     registers as bf16 mma.sync A fragments; h_t (forward) / partial dL/dh
//     (backward) cross the cluster as st.async + mbarrier complete_tx.
// What is different (profiles/r1_rnn_stalls.md): in the one-role kernels a
// third of every step was the compute warp waiting on its OWN global-memory
// instructions -- write-after-read scoreboard stalls on the operands of 8
// cp.async, 8 st.global and 8 st.async per thread, constant-bank reloads and
// 64-bit address arithmetic -- all in the single instruction stream that also
// carries the step's dependency chain.  Here
//   * the projection / reserve tensors are UNIT-MAJOR ([T][N][H][G]: the G
//     gates of a cell are adjacent), so a CTA's slice of a chunk is one
//     contiguous run (U*G*4 bytes) and a thread's values are one 16-byte word;
//   * a dedicated I/O warp streams those runs: inputs global -> shared ring
//     (cp.async, completion signalled by cp.async.mbarrier.arrive), outputs
//     shared staging ring -> global (coalesced 16-byte stores);
//   * the compute warps touch shared memory only: per step they wait for
//     h_{t-1}, run ldmatrix + HMMA + the gate math, send h_t with st.async,
//     drop their outputs into the staging slot and pick up the next inputs.
// Ring protocol (kS = 4 slots, step s uses slot s % 4 of every ring):
//   ifull[slot]  I/O warp -> compute: inputs of the step have landed
//   ofull[slot]  compute -> I/O warp: outputs of the step are staged, and the
//                step's inputs have been consumed (the slot can be refilled
//                with step s + 4)
// The I/O warp refills slot s only after copying out the outputs of step s,
// so "inputs of step s + 4 have landed" also tells the compute warps that the
// output slot of step s is free again.
//
// Sizes: the cluster size CL is a template parameter (8 for hidden sizes that are multiples
// of 64, 4 for the odd multiples of 32 up to 224), U = H/CL units per CTA, U/8 compute
// warps.  The W_hh slice is register-resident up to 128 registers per thread; beyond that
// (H > 256 forward, G*U > 128 gate rows backward) the remaining k tiles are kept in shared
// memory in fragment order -- each lane re-reads the 16 bytes it wrote, one conflict-free
// LDS.128 per HMMA -- so sizes 320 / 384 / 448 run on the same kernels.
#pragma once
#include "rnn_common.cuh"

namespace ty {

struct RnnWsArgs {
    const float *w_hh;        // [G*H][H] fp32, gate-major (the parameter itself)
    int T, N, reverse;
    // forward
    const float *xproj;       // [T][N][H][G]
    const float *bias;        // [G*H] gate-major, may be null
    float *y;                 // [T][N][H]
    __nv_bfloat16 *y16;       // [T][N][H], may be null
    float *gates;             // [T][N][H][4]   LSTM: i f g o    GRU: r z n (W_hn h)
    float *cstate;            // [T][N][H]      LSTM only
    // backward
    const float *dy;          // [T][N][H]
    __nv_bfloat16 *dx16;      // [T][N][H][G]   gradient of xproj
    __nv_bfloat16 *dhid16;    // GRU: [T][N][H][3] (dr, dz, d(W_hn h)): gradient of the hidden-side product
    float *dbias;             // [G*H] gate-major, += ; may be null
};

constexpr int kS = 4;         // ring slots
// Bytes per partial dL/dh sum crossing the cluster in the backward kernel.  2 = bf16 pairs
// (default): the reduce-scatter then moves 512 B per CTA pair and step, like the forward exchange,
// instead of 1 KB -- backward kernel 0.435 ms per layer against 0.510 ms (config A, inside the
// step), and against the fp32 parity mode over 800 steps the gradients are where they were
// (input gradient 0.38 % relative, weight gradients 0.25-0.42 %): the partial sums are products of
// bf16 operands whose rounding already sets that level.  4 = fp32 (-DTY_BWD_FP32_PARTIALS).
#ifdef TY_BWD_FP32_PARTIALS
constexpr int kPartialBytes = 4;
#else
constexpr int kPartialBytes = 2;
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the executing thread's earlier cp.async operations arrive on `bar` when they complete
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int pinned_tid_x() {   // not rematerialised as S2R inside the step loop
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
template <int CELL, int H, int CL>
struct FwdLayout {
    static constexpr int G = Cell<CELL>::G;
    static constexpr int U = H / CL;                        // units per CTA
    static constexpr int NCW = U / 8;                       // compute warps (8 units each)
    static constexpr int NCT = NCW * 32;
    static constexpr int KT = H / 16;                       // k tiles of the recurrent product
    static constexpr int KR = KT < 16 ? KT : 16;            // ... held in registers
    static constexpr int KS = KT - KR;                      // ... held in shared memory
    static_assert(H % CL == 0 && U % 8 == 0 && KT % 2 == 0, "hidden size / cluster size");
    static constexpr int XROW = U * G + (G == 3 ? 4 : 0);   // floats per chunk row of an x slot
    static constexpr int CROW = U + 4;                      // floats per chunk row of the c / y stages
    static constexpr int YROW = U + 8;                      // bf16 per chunk row of the y16 stage
    static constexpr int HS_BYTES = 2 * H * kNB * 2;
    static constexpr int XS_BYTES = kNB * XROW * 4;         // one slot
    static constexpr int OG_BYTES = kNB * U * 16;
    static constexpr int OC_BYTES = kNB * CROW * 4;
    static constexpr int OY16_BYTES = kNB * YROW * 2;
    static constexpr int off_hs = 0;
    static constexpr int off_xs = off_hs + HS_BYTES;
    static constexpr int off_og = off_xs + kS * XS_BYTES;
    static constexpr int off_oc = off_og + kS * OG_BYTES;
    static constexpr int off_oy = off_oc + kS * OC_BYTES;
    static constexpr int off_oy16 = off_oy + kS * OC_BYTES;
    static constexpr int off_bar = off_oy16 + kS * OY16_BYTES;
    static constexpr int off_as = (off_bar + (2 + 2 * kS) * 8 + 15) / 16 * 16;   // A fragments of k tiles >= KR
    static constexpr int total = off_as + NCW * 2 * KS * 512;
    static_assert(XS_BYTES % 16 == 0 && OC_BYTES % 16 == 0 && OY16_BYTES % 16 == 0, "alignment");
    static_assert(total <= 227 * 1024, "shared memory");
};

template <int CELL, int H, int CL>
__global__ void __launch_bounds__(FwdLayout<CELL, H, CL>::NCT + 32, 1)
    rnn_ws_forward_kernel(const RnnWsArgs a) {
    using L = FwdLayout<CELL, H, CL>;
    constexpr int G = L::G, U = L::U, KT = L::KT, KR = L::KR, KS = L::KS, NCW = L::NCW, NCT = L::NCT;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + L::off_bar);   // [2]  h_{t-1} has arrived
    uint64_t *const xfull = full + 2;                                         // [kS]
    uint64_t *const ofull = xfull + kS;                                       // [kS]

    const int tid = pinned_tid_x(), lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / CL;
    const int T = a.T, N = a.N;

    // h tiles and the x ring start at zero: the columns of chunks >= N are never
    // loaded and must stay finite
    for (int i = tid; i < L::off_og / 4; i += NCT + 32) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        for (int i = 0; i < kS; i++) {
            mbar_init(&xfull[i], 32);
            mbar_init(&ofull[i], NCT);
        }
        fence_mbar_init();
    }

    // --- W_hh slice -> A fragments (compute warps; registers, whole sequence) ---
    const int r = lane >> 2, q = lane & 3;
    const int ul = (warp < NCW ? warp : 0) * 8 + r;      // local unit of this thread's cells
    const int unit = rank * U + ul;
    uint32_t A[2][KR][4];
    // fragments of k tiles >= KR: [warp][m][kt - KR][lane] 16-byte words, written and read by
    // the same lane
    uint4 *const As = reinterpret_cast<uint4 *>(smem + L::off_as) + (size_t)(warp < NCW ? warp : 0) * 2 * KS * 32 + lane;
    if (warp < NCW) {
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const int glo = 2 * m, ghi = 2 * m + 1;
            const float *wlo = a.w_hh + ((size_t)glo * H + unit) * H;
            const float *whi = a.w_hh + ((size_t)(ghi < G ? ghi : 0) * H + unit) * H;
#pragma unroll
            for (int kt = 0; kt < KT; kt++) {
                const int k = 16 * kt + 2 * q;
                uint32_t f[4];
                f[0] = pack_bf16(wlo[k], wlo[k + 1]);
                f[2] = pack_bf16(wlo[k + 8], wlo[k + 9]);
                if (ghi < G) {
                    f[1] = pack_bf16(whi[k], whi[k + 1]);
                    f[3] = pack_bf16(whi[k + 8], whi[k + 9]);
                } else {
                    f[1] = 0u;
                    f[3] = 0u;
                }
                if (kt < KR) {
#pragma unroll
                    for (int e = 0; e < 4; e++) A[m][kt < KR ? kt : 0][e] = f[e];
                } else {
                    As[(m * KS + (kt - KR)) * 32] = make_uint4(f[0], f[1], f[2], f[3]);
                }
            }
        }
    }
    cluster_sync_all();

    if (warp == NCW) {
        // ===================== I/O warp =====================
        const int nvalid = min(kNB, N - group * kNB);        // chunks of this group that exist
        auto cell0 = [&](int s) {                            // first cell of (time of step s, chunk 0 of the group, unit 0 of the CTA)
            const int t = a.reverse ? T - 1 - s : s;
            return ((size_t)t * N + (size_t)group * kNB) * H + (size_t)rank * U;
        };
        constexpr int PX = U * G / 4;                        // 16-byte pieces of a chunk's x run
        auto issue_x = [&](int s, int slot) {
            const float *src = a.xproj + cell0(s) * G;
            unsigned char *dst = smem + L::off_xs + slot * L::XS_BYTES;
#pragma unroll
            for (int i0 = 0; i0 < kNB * PX; i0 += 32) {
                const int i = i0 + lane;
                const int n = i / PX, p = i - n * PX;
                if (i < kNB * PX && n < nvalid)
                    cp_async16(dst + (n * L::XROW + p * 4) * 4, src + (size_t)n * H * G + p * 4);
            }
            cp_async_mbar_arrive_noinc(&xfull[slot]);
        };
        for (int s = 0; s < kS && s < T; s++) issue_x(s, s);
        for (int j = 0; j < T; j++) {
            const int slot = j & (kS - 1);
            mbar_wait(&ofull[slot], (uint32_t)(j >> 2) & 1u);
            const size_t c0 = cell0(j);
            // ---- gates: kNB runs of U * 16 bytes ----
            {
                constexpr int NP = kNB * U;
                constexpr int IT = (NP + 31) / 32;
                const float4 *src = reinterpret_cast<const float4 *>(smem + L::off_og + slot * L::OG_BYTES);
                float4 *dst = reinterpret_cast<float4 *>(a.gates) + c0;
                float4 v[IT];
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    const int i = k * 32 + lane;
                    if (i < NP) v[k] = src[i];
                }
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    const int i = k * 32 + lane;
                    const int n = i / U, p = i - n * U;
                    if (i < NP && n < nvalid) __stcs(dst + (size_t)n * H + p, v[k]);
                }
            }
            // ---- cell state (LSTM) and y: kNB runs of U * 4 bytes ----
            {
                constexpr int PP = U / 4;
                constexpr int NP = kNB * PP;
                constexpr int IT = (NP + 31) / 32;
                const unsigned char *sc = smem + L::off_oc + slot * L::OC_BYTES;
                const unsigned char *sy = smem + L::off_oy + slot * L::OC_BYTES;
                float4 vc[IT], vy[IT];
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    const int i = k * 32 + lane;
                    const int n = i / PP, p = i - n * PP;
                    if (i < NP) {
                        if (CELL == kLstm)
                            vc[k] = *reinterpret_cast<const float4 *>(sc + (n * L::CROW + p * 4) * 4);
                        vy[k] = *reinterpret_cast<const float4 *>(sy + (n * L::CROW + p * 4) * 4);
                    }
                }
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    const int i = k * 32 + lane;
                    const int n = i / PP, p = i - n * PP;
                    if (i < NP && n < nvalid) {
                        const size_t o = c0 + (size_t)n * H + p * 4;
                        if (CELL == kLstm) __stcs(reinterpret_cast<float4 *>(a.cstate + o), vc[k]);
                        *reinterpret_cast<float4 *>(a.y + o) = vy[k];
                    }
                }
            }
            // ---- bf16 copy of y: kNB runs of U * 2 bytes ----
            if (a.y16) {
                constexpr int PP = U / 8;
                constexpr int NP = kNB * PP;
                const unsigned char *sy = smem + L::off_oy16 + slot * L::OY16_BYTES;
#pragma unroll
                for (int i0 = 0; i0 < NP; i0 += 32) {
                    const int i = i0 + lane;
                    const int n = i / PP, p = i - n * PP;
                    if (i < NP && n < nvalid) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(sy + (n * L::YROW + p * 8) * 2);
                        *reinterpret_cast<uint4 *>(a.y16 + c0 + (size_t)n * H + p * 8) = v;
                    }
                }
            }
            if (j + kS < T) issue_x(j + kS, slot);
        }
    } else {
        // ===================== compute warps =====================
        float cst[2] = {0.f, 0.f};    // LSTM cell state / GRU previous h (fp32)
        uint32_t phase = 0u;          // bit b = parity to wait for on full[b]
        float bias[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.bias) {
#pragma unroll
            for (int g = 0; g < G; g++) bias[g] = a.bias[(size_t)g * H + unit];
        }
        const uint32_t hs_base = smem_u32(smem + L::off_hs);
        const uint32_t bar_base = smem_u32(&full[0]);
        const uint32_t ld_off = (uint32_t)(lane * 16);
        const uint32_t send_off = (uint32_t)((unit * kNB + 2 * q) * 2);

        float xp[4][2];               // projection values of the step about to run
        auto load_x = [&](int slot) {
            const unsigned char *xs = smem + L::off_xs + slot * L::XS_BYTES;
#pragma unroll
            for (int col = 0; col < 2; col++) {
                const float *p = reinterpret_cast<const float *>(xs) + (2 * q + col) * L::XROW + ul * G;
                if (G == 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(p);
                    xp[0][col] = v.x; xp[1][col] = v.y; xp[2][col] = v.z; xp[3][col] = v.w;
                } else {
                    xp[0][col] = p[0]; xp[1][col] = p[1]; xp[2][col] = p[2]; xp[3][col] = 0.f;
                }
            }
        };

        auto step = [&](const int s, auto slot_c) {
            constexpr int SLOT = decltype(slot_c)::value;
            constexpr int cur = SLOT & 1, nxt = cur ^ 1;
            if (tid == 0 && s + 1 < T)      // arm the barrier that collects h_t
                mbar_arrive_expect_tx(&full[nxt], CL * kNB * U * 2);
            float acc[2][4][4];
#pragma unroll
            for (int m = 0; m < 2; m++)
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[m][c][e] = 0.f;
            // element e of tile m: gate 2m + (e >> 1), column e & 1.  Chain 0 is seeded with
            // the bias, chain 1 with the projection (GRU: the x side of the n gate must stay
            // outside W_hn h; it is added below).
            acc[0][0][0] = bias[0]; acc[0][0][1] = bias[0];
            acc[0][0][2] = bias[1]; acc[0][0][3] = bias[1];
            acc[0][1][0] = xp[0][0]; acc[0][1][1] = xp[0][1];
            acc[0][1][2] = xp[1][0]; acc[0][1][3] = xp[1][1];
            if (CELL == kLstm) {
                acc[1][0][0] = bias[2]; acc[1][0][1] = bias[2];
                acc[1][0][2] = bias[3]; acc[1][0][3] = bias[3];
                acc[1][1][0] = xp[2][0]; acc[1][1][1] = xp[2][1];
                acc[1][1][2] = xp[3][0]; acc[1][1][3] = xp[3][1];
            }
            if (s > 0) {
                mbar_wait(&full[cur], (phase >> cur) & 1u);
                phase ^= 1u << cur;
            }
            const uint32_t hcur = hs_base + (uint32_t)(cur * kNB * H * 2) + ld_off;
#pragma unroll
            for (int kp = 0; kp < KR / 2; kp++) {
                uint32_t bf[4];
                ldmatrix_x4_trans(bf, hcur + kp * 512);
                mma_bf16(acc[0][(2 * kp) & 3], A[0][2 * kp], bf[0], bf[1]);
                mma_bf16(acc[1][(2 * kp) & 3], A[1][2 * kp], bf[0], bf[1]);
                mma_bf16(acc[0][(2 * kp + 1) & 3], A[0][2 * kp + 1], bf[2], bf[3]);
                mma_bf16(acc[1][(2 * kp + 1) & 3], A[1][2 * kp + 1], bf[2], bf[3]);
            }
            if constexpr (KS > 0) {
#pragma unroll
                for (int kp = KR / 2; kp < KT / 2; kp++) {
                    uint32_t bf[4];
                    ldmatrix_x4_trans(bf, hcur + kp * 512);
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int ks = 2 * kp + half - KR;
#pragma unroll
                        for (int m = 0; m < 2; m++) {
                            const uint4 w = As[(m * KS + ks) * 32];
                            const uint32_t f[4] = {w.x, w.y, w.z, w.w};
                            mma_bf16(acc[m][(2 * kp + half) & 3], f, bf[2 * half], bf[2 * half + 1]);
                        }
                    }
                }
            }
            float pre[2][4];
#pragma unroll
            for (int m = 0; m < 2; m++)
#pragma unroll
                for (int e = 0; e < 4; e++)
                    pre[m][e] = (acc[m][0][e] + acc[m][1][e]) + (acc[m][2][e] + acc[m][3][e]);

            float hnew[2], cnew[2];
            float4 sv[2];
#pragma unroll
            for (int col = 0; col < 2; col++) {
                if (CELL == kLstm) {
                    const float gi = sigmoidf_(pre[0][col]);
                    const float gf = sigmoidf_(pre[0][2 + col]);
                    const float gg = tanhf_(pre[1][col]);
                    const float go = sigmoidf_(pre[1][2 + col]);
                    const float c = gf * cst[col] + gi * gg;
                    cst[col] = c;
                    cnew[col] = c;
                    hnew[col] = go * tanhf_(c);
                    sv[col] = make_float4(gi, gf, gg, go);
                } else {
                    const float hn = pre[1][col];
                    const float gr = sigmoidf_(pre[0][col]);
                    const float gz = sigmoidf_(pre[0][2 + col]);
                    const float gn = tanhf_(xp[2][col] + bias[2] + gr * hn);
                    const float h = (1.0f - gz) * gn + gz * cst[col];
                    cst[col] = h;
                    cnew[col] = 0.f;
                    hnew[col] = h;
                    sv[col] = make_float4(gr, gz, gn, hn);
                }
            }
            const uint32_t pair = pack_bf16(hnew[0], hnew[1]);
            if (s + 1 < T) {
                // Exchange h_t: one bf16 pair (this unit, chunks 2q and 2q+1) stored straight
                // from registers into the unit-major tile of all 8 CTAs of the cluster (own
                // included); the receivers' mbarriers count the bytes.
                const uint32_t dst = hs_base + (uint32_t)(nxt * kNB * H * 2) + send_off;
                const uint32_t bar = bar_base + (uint32_t)(nxt * 8);
#pragma unroll
                for (uint32_t peer = 0; peer < CL; peer++)
                    st_async_b32(mapa(dst, peer), pair, mapa(bar, peer));
            }
            // ---- outputs of the step -> staging slot (the I/O warp writes them out) ----
            {
                unsigned char *og = smem + L::off_og + SLOT * L::OG_BYTES;
                float *oc = reinterpret_cast<float *>(smem + L::off_oc + SLOT * L::OC_BYTES);
                float *oy = reinterpret_cast<float *>(smem + L::off_oy + SLOT * L::OC_BYTES);
                __nv_bfloat16 *oy16 = reinterpret_cast<__nv_bfloat16 *>(smem + L::off_oy16 + SLOT * L::OY16_BYTES);
#pragma unroll
                for (int col = 0; col < 2; col++) {
                    const int n = 2 * q + col;
                    *reinterpret_cast<float4 *>(og + (n * U + ul) * 16) = sv[col];
                    if (CELL == kLstm) oc[n * L::CROW + ul] = cnew[col];
                    oy[n * L::CROW + ul] = hnew[col];
                }
                reinterpret_cast<unsigned short *>(oy16)[(2 * q) * L::YROW + ul] = (unsigned short)(pair & 0xffffu);
                reinterpret_cast<unsigned short *>(oy16)[(2 * q + 1) * L::YROW + ul] = (unsigned short)(pair >> 16);
            }
            mbar_arrive(&ofull[SLOT]);
            // ---- projection values of the next step ----
            if (s + 1 < T) {
                mbar_wait(&xfull[(SLOT + 1) & (kS - 1)], (uint32_t)((s + 1) >> 2) & 1u);
                load_x((SLOT + 1) & (kS - 1));
            }
        };

        mbar_wait(&xfull[0], 0u);
        load_x(0);
        using std::integral_constant;
        int s = 0;
        for (; s + 3 < T; s += 4) {
            step(s, integral_constant<int, 0>{});
            step(s + 1, integral_constant<int, 1>{});
            step(s + 2, integral_constant<int, 2>{});
            step(s + 3, integral_constant<int, 3>{});
        }
        if (s < T) step(s, integral_constant<int, 0>{});
        if (s + 1 < T) step(s + 1, integral_constant<int, 1>{});
        if (s + 2 < T) step(s + 2, integral_constant<int, 2>{});
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Backward (BPTT).  K of the per-step product is the CTA's G*U gate rows in
// unit-major order (k = local unit * G + gate, padded to a multiple of 16), M is
// all H hidden units; partial dL/dh_{t-1} tiles are reduce-scattered to their
// owners through DSMEM.  The bf16 gate gradients are written once, into a ring
// slot that is both the B operand of the product and the staging buffer of the
// dxproj output (LSTM) / the hidden-side gradient (GRU).
template <int CELL, int H, int CL>
struct BwdLayout {
    static constexpr int G = Cell<CELL>::G;
    static constexpr int U = H / CL;
    static constexpr int NCW = U / 8;
    static constexpr int NCT = NCW * 32;
    static constexpr int MT = CL / 2;                       // m tiles (16 units) per warp: (H/16) / NCW
    static constexpr int KL = (G * U + 15) / 16 * 16;       // local gate rows, padded
    static constexpr int KT = KL / 16;
    static constexpr int KR = KT < 32 / MT ? KT : 32 / MT;  // k tiles in registers (<= 128 registers)
    static constexpr int KS = KT - KR;                      // k tiles in shared memory
    static_assert(H % CL == 0 && U % 8 == 0 && CL % 2 == 0 && (H / 16) % NCW == 0, "hidden size / cluster size");
    static constexpr int DS = KL + 8;                       // bf16 per chunk row of a ds slot
    static constexpr int CROW = U + 4;                      // floats per chunk row of dy / c / prev
    static constexpr int RS_BYTES = 2 * CL * U * kNB * kPartialBytes;
    static constexpr int DS_BYTES = kNB * DS * 2;           // one slot
    static constexpr int IG_BYTES = kNB * U * 16;
    static constexpr int IC_BYTES = kNB * CROW * 4;
    static constexpr int off_rs = 0;
    static constexpr int off_ds = off_rs + RS_BYTES;
    static constexpr int off_ox = off_ds + kS * DS_BYTES;                       // GRU: x-side gradient stage (same row format as ds)
    static constexpr int off_ig = off_ox + (CELL == kGru ? kS * DS_BYTES : 0);
    static constexpr int off_idy = off_ig + kS * IG_BYTES;
    static constexpr int off_ic = off_idy + kS * IC_BYTES;                      // LSTM: c_t
    static constexpr int off_ip = off_ic + (CELL == kLstm ? kS * IC_BYTES : 0); // c_{t-1} / h_{t-1}
    static constexpr int off_bar = off_ip + kS * IC_BYTES;
    static constexpr int off_as = (off_bar + (2 + 2 * kS) * 8 + 15) / 16 * 16;   // A fragments of k tiles >= KR
    static constexpr int total = off_as + NCW * MT * KS * 512;
    static_assert(DS_BYTES % 16 == 0 && IC_BYTES % 16 == 0 && (DS * 2) % 16 == 0, "alignment");
    static_assert(total <= 227 * 1024, "shared memory");
};

template <int CELL, int H, int CL>
__global__ void __launch_bounds__(BwdLayout<CELL, H, CL>::NCT + 32, 1)
    rnn_ws_backward_kernel(const RnnWsArgs a) {
    using L = BwdLayout<CELL, H, CL>;
    constexpr int G = L::G, U = L::U, KT = L::KT, KR = L::KR, KS = L::KS, DS = L::DS, NCW = L::NCW, NCT = L::NCT;
    constexpr int MT = L::MT;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + L::off_bar);   // [2] partial sums have arrived
    uint64_t *const ifull = full + 2;                                         // [kS]
    uint64_t *const ofull = ifull + kS;                                       // [kS]

    const int tid = pinned_tid_x(), lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / CL;
    const int T = a.T, N = a.N;

    // everything starts at zero: pad columns of ds, slots of chunks >= N
    for (int i = tid; i < L::off_bar / 4; i += NCT + 32) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        for (int i = 0; i < kS; i++) {
            mbar_init(&ifull[i], 32);
            mbar_init(&ofull[i], NCT);
        }
        fence_mbar_init();
    }

    const int r = lane >> 2, q = lane & 3;
    const int cw = warp < NCW ? warp : 0;
    const int ul = cw * 8 + r;                   // local unit of this thread's cells
    const int unit = rank * U + ul;
    // --- W_hh^T slice -> A fragments: A[m = hidden unit][k = local gate row ul*G + g] ---
    uint32_t A[MT][KR][4];
    uint4 *const As = reinterpret_cast<uint4 *>(smem + L::off_as) + (size_t)cw * MT * KS * 32 + lane;
    if (warp < NCW) {
        auto wrow = [&](int kl, int h) -> float {
            if (kl >= G * U) return 0.f;
            const int u = kl / G, g = kl - u * G;
            return a.w_hh[((size_t)g * H + rank * U + u) * H + h];
        };
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int h0 = 16 * MT * cw + 16 * mt + r;
#pragma unroll
            for (int kt = 0; kt < KT; kt++) {
                const int k = 16 * kt + 2 * q;
                uint32_t f[4];
                f[0] = pack_bf16(wrow(k, h0), wrow(k + 1, h0));
                f[1] = pack_bf16(wrow(k, h0 + 8), wrow(k + 1, h0 + 8));
                f[2] = pack_bf16(wrow(k + 8, h0), wrow(k + 9, h0));
                f[3] = pack_bf16(wrow(k + 8, h0 + 8), wrow(k + 9, h0 + 8));
                if (kt < KR) {
#pragma unroll
                    for (int e = 0; e < 4; e++) A[mt][kt < KR ? kt : 0][e] = f[e];
                } else {
                    As[(mt * KS + (kt - KR)) * 32] = make_uint4(f[0], f[1], f[2], f[3]);
                }
            }
        }
    }
    cluster_sync_all();

    auto tindex = [&](int sf) { return a.reverse ? T - 1 - sf : sf; };   // forward step -> time

    if (warp == NCW) {
        // ===================== I/O warp =====================
        const int nvalid = min(kNB, N - group * kNB);
        auto cell0 = [&](int sf) {
            return ((size_t)tindex(sf) * N + (size_t)group * kNB) * H + (size_t)rank * U;
        };
        // inputs of backward step s (forward step sf = T-1-s) -> ring slot
        auto issue_in = [&](int s, int slot) {
            const int sf = T - 1 - s;
            const size_t c0 = cell0(sf);
            {   // saved gates
                constexpr int NP = kNB * U;
                const float4 *src = reinterpret_cast<const float4 *>(a.gates) + c0;
                unsigned char *dst = smem + L::off_ig + slot * L::IG_BYTES;
#pragma unroll
                for (int i0 = 0; i0 < NP; i0 += 32) {
                    const int i = i0 + lane;
                    const int n = i / U, p = i - n * U;
                    if (i < NP && n < nvalid) cp_async16(dst + i * 16, src + (size_t)n * H + p);
                }
            }
            {   // dy, c_t (LSTM), c_{t-1} (LSTM) / h_{t-1} (GRU)
                constexpr int PP = U / 4;
                constexpr int NP = kNB * PP;
                const float *prev_src = nullptr;
                if (sf > 0) prev_src = (CELL == kLstm ? a.cstate : a.y) + cell0(sf - 1);
                unsigned char *ddy = smem + L::off_idy + slot * L::IC_BYTES;
                unsigned char *dc = smem + L::off_ic + slot * L::IC_BYTES;
                unsigned char *dp = smem + L::off_ip + slot * L::IC_BYTES;
#pragma unroll
                for (int i0 = 0; i0 < NP; i0 += 32) {
                    const int i = i0 + lane;
                    const int n = i / PP, p = i - n * PP;
                    if (i < NP && n < nvalid) {
                        const int so = (n * L::CROW + p * 4) * 4;
                        const size_t go = (size_t)n * H + p * 4;
                        cp_async16(ddy + so, a.dy + c0 + go);
                        if (CELL == kLstm) cp_async16(dc + so, a.cstate + c0 + go);
                        if (prev_src) cp_async16(dp + so, prev_src + go);
                    }
                }
            }
            cp_async_mbar_arrive_noinc(&ifull[slot]);
        };
        for (int s = 0; s < kS && s < T; s++) issue_in(s, s);
        for (int j = 0; j < T; j++) {
            const int slot = j & (kS - 1);
            mbar_wait(&ofull[slot], (uint32_t)(j >> 2) & 1u);
            const size_t c0 = cell0(T - 1 - j);
            // gate gradients: kNB runs of U*G bf16
            constexpr int PP = U * G / 8;              // 16-byte pieces per run
            constexpr int NP = kNB * PP;
            constexpr int IT = (NP + 31) / 32;
            const unsigned char *sd = smem + L::off_ds + slot * L::DS_BYTES;
            const unsigned char *sx = smem + L::off_ox + slot * L::DS_BYTES;
            uint4 vd[IT], vx[IT];
#pragma unroll
            for (int k = 0; k < IT; k++) {
                const int i = k * 32 + lane;
                const int n = i / PP, p = i - n * PP;
                if (i < NP) {
                    vd[k] = *reinterpret_cast<const uint4 *>(sd + n * DS * 2 + p * 16);
                    if (CELL == kGru) vx[k] = *reinterpret_cast<const uint4 *>(sx + n * DS * 2 + p * 16);
                }
            }
#pragma unroll
            for (int k = 0; k < IT; k++) {
                const int i = k * 32 + lane;
                const int n = i / PP, p = i - n * PP;
                if (i < NP && n < nvalid) {
                    const size_t o = (c0 + (size_t)n * H) * G + p * 8;
                    if (CELL == kLstm) {
                        *reinterpret_cast<uint4 *>(a.dx16 + o) = vd[k];
                    } else {
                        *reinterpret_cast<uint4 *>(a.dhid16 + o) = vd[k];
                        *reinterpret_cast<uint4 *>(a.dx16 + o) = vx[k];
                    }
                }
            }
            if (j + kS < T) issue_in(j + kS, slot);
        }
    } else {
        // ===================== compute warps =====================
        uint32_t phase = 0u;
        float carry[2] = {0.f, 0.f};   // LSTM: dL/dc carried back; GRU: z * dL/dh carried back
        float dbacc[G];                // bias gradient of this thread's unit
#pragma unroll
        for (int g = 0; g < G; g++) dbacc[g] = 0.f;
        const uint32_t ld_off = (uint32_t)(((lane & 7) * DS + 8 * (lane >> 3)) * 2);
        const uint32_t rs_base = smem_u32(smem + L::off_rs);
        const uint32_t bar_base = smem_u32(&full[0]);
        const float *rs = reinterpret_cast<const float *>(smem + L::off_rs);

        float4 gt[2];
        float dyv[2], cv[2], pv[2];
        auto load_in = [&](int slot) {
#pragma unroll
            for (int col = 0; col < 2; col++) {
                const int n = 2 * q + col;
                gt[col] = *reinterpret_cast<const float4 *>(smem + L::off_ig + slot * L::IG_BYTES + (n * U + ul) * 16);
                const int so = slot * L::IC_BYTES + (n * L::CROW + ul) * 4;
                dyv[col] = *reinterpret_cast<const float *>(smem + L::off_idy + so);
                cv[col] = CELL == kLstm ? *reinterpret_cast<const float *>(smem + L::off_ic + so) : 0.f;
                pv[col] = *reinterpret_cast<const float *>(smem + L::off_ip + so);
            }
        };

        auto step = [&](const int s, auto slot_c) {
            constexpr int SLOT = decltype(slot_c)::value;
            constexpr int cur = SLOT & 1, nxt = cur ^ 1;
            const int sf = T - 1 - s;
            if (tid == 0 && s + 1 < T) mbar_arrive_expect_tx(&full[nxt], CL * U * kNB * kPartialBytes);
            float dh[2] = {dyv[0], dyv[1]};
            if (s > 0) {
                mbar_wait(&full[cur], (phase >> cur) & 1u);
                phase ^= 1u << cur;
                float sx = 0.f, sy = 0.f;
#pragma unroll
                for (int j = 0; j < CL; j++) {
                    if (kPartialBytes == 4) {
                        const float2 v = *reinterpret_cast<const float2 *>(
                            rs + ((cur * CL + j) * U + ul) * kNB + 2 * q);
                        sx += v.x; sy += v.y;
                    } else {
                        const uint32_t w = *reinterpret_cast<const uint32_t *>(
                            reinterpret_cast<const unsigned char *>(rs) + (((cur * CL + j) * U + ul) * kNB + 2 * q) * 2);
                        sx += __uint_as_float(w << 16); sy += __uint_as_float(w & 0xffff0000u);
                    }
                }
                dh[0] += sx; dh[1] += sy;
            }
            unsigned char *sd = smem + L::off_ds + SLOT * L::DS_BYTES;
            unsigned char *sxo = smem + L::off_ox + SLOT * L::DS_BYTES;
#pragma unroll
            for (int col = 0; col < 2; col++) {
                const int n = 2 * q + col;
                const float prev = sf > 0 ? pv[col] : 0.f;
                if (CELL == kLstm) {
                    const float gi = gt[col].x, gf = gt[col].y, gg = gt[col].z, go = gt[col].w;
                    const float tc = tanhf_(cv[col]);
                    const float d = dh[col];
                    const float dc = carry[col] + d * go * (1.0f - tc * tc);
                    const float d3 = d * tc * go * (1.0f - go);
                    const float d0 = dc * gg * gi * (1.0f - gi);
                    const float d2 = dc * gi * (1.0f - gg * gg);
                    const float d1 = dc * prev * gf * (1.0f - gf);
                    carry[col] = dc * gf;
                    dbacc[0] += d0; dbacc[1] += d1; dbacc[2] += d2; dbacc[3] += d3;
                    uint2 w;
                    w.x = pack_bf16(d0, d1);
                    w.y = pack_bf16(d2, d3);
                    *reinterpret_cast<uint2 *>(sd + (n * DS + ul * 4) * 2) = w;
                } else {
                    const float gr = gt[col].x, gz = gt[col].y, gn = gt[col].z, hn = gt[col].w;
                    const float d = dh[col] + carry[col];
                    const float dn = d * (1.0f - gz) * (1.0f - gn * gn);      // d n_pre (x side)
                    const float d1 = d * (prev - gn) * gz * (1.0f - gz);
                    const float d0 = dn * hn * gr * (1.0f - gr);
                    carry[col] = d * gz;
                    const float dhn = dn * gr;                                // hidden-side n gradient
                    dbacc[0] += d0; dbacc[1] += d1; dbacc[2] += dn;
                    __nv_bfloat16 *pd = reinterpret_cast<__nv_bfloat16 *>(sd) + n * DS + ul * 3;
                    __nv_bfloat16 *px = reinterpret_cast<__nv_bfloat16 *>(sxo) + n * DS + ul * 3;
                    const __nv_bfloat16 b0 = __float2bfloat16(d0), b1 = __float2bfloat16(d1);
                    pd[0] = b0; pd[1] = b1; pd[2] = __float2bfloat16(dhn);
                    px[0] = b0; px[1] = b1; px[2] = __float2bfloat16(dn);
                }
            }
            mbar_arrive(&ofull[SLOT]);
            if (s + 1 < T) {
                named_bar_sync(1, NCT);          // every compute warp's gate gradients are in the slot
                float acc[MT][2][4];
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int c = 0; c < 2; c++)
#pragma unroll
                        for (int e = 0; e < 4; e++) acc[mt][c][e] = 0.f;
                const uint32_t ds_base = smem_u32(sd) + ld_off;
#pragma unroll
                for (int kp = 0; kp < (KT + 1) / 2; kp++) {
                    // (odd number of k tiles: the upper half of the last x4 is padding, not multiplied)
                    uint32_t bf[4];
                    ldmatrix_x4(bf, ds_base + kp * 64);
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int kt = 2 * kp + half;
                        if (kt < KT) {
#pragma unroll
                            for (int mt = 0; mt < MT; mt++) {
                                if (kt < KR) {
                                    mma_bf16(acc[mt][half], A[mt][kt < KR ? kt : 0], bf[2 * half], bf[2 * half + 1]);
                                } else {
                                    const uint4 w = As[(mt * KS + (kt - KR)) * 32];
                                    const uint32_t f[4] = {w.x, w.y, w.z, w.w};
                                    mma_bf16(acc[mt][half], f, bf[2 * half], bf[2 * half + 1]);
                                }
                            }
                        }
                    }
                }
                // reduce-scatter: rows of this warp's tiles belong to the CTA owning those units
                const uint32_t rs_nxt = rs_base + (uint32_t)(((nxt * CL + rank) * U * kNB) * kPartialBytes);
                const uint32_t bar = bar_base + (uint32_t)(nxt * 8);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int h = 16 * MT * cw + 16 * mt + 8 * half + r;
                        const uint32_t dest = h / U;
                        const int hl = h - dest * U;
                        const float x = acc[mt][0][2 * half] + acc[mt][1][2 * half];
                        const float y = acc[mt][0][2 * half + 1] + acc[mt][1][2 * half + 1];
                        const uint32_t local = rs_nxt + (uint32_t)((hl * kNB + 2 * q) * kPartialBytes);
                        if (kPartialBytes == 4) st_async_v2(mapa(local, dest), x, y, mapa(bar, dest));
                        else st_async_b32(mapa(local, dest), pack_bf16(x, y), mapa(bar, dest));
                    }
                }
                mbar_wait(&ifull[(SLOT + 1) & (kS - 1)], (uint32_t)((s + 1) >> 2) & 1u);
                load_in((SLOT + 1) & (kS - 1));
            }
        };

        mbar_wait(&ifull[0], 0u);
        load_in(0);
        using std::integral_constant;
        int s = 0;
        for (; s + 3 < T; s += 4) {
            step(s, integral_constant<int, 0>{});
            step(s + 1, integral_constant<int, 1>{});
            step(s + 2, integral_constant<int, 2>{});
            step(s + 3, integral_constant<int, 3>{});
        }
        if (s < T) step(s, integral_constant<int, 0>{});
        if (s + 1 < T) step(s + 1, integral_constant<int, 1>{});
        if (s + 2 < T) step(s + 2, integral_constant<int, 2>{});
        if (a.dbias) {
            // chunks >= N contribute zeros: their slots are never loaded
#pragma unroll
            for (int g = 0; g < G; g++) atomicAdd(a.dbias + (size_t)g * H + unit, dbacc[g]);
        }
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Launch: one cluster of CL CTAs per group of kNB chunks (cluster size as a launch attribute).
template <typename K>
static int launch_ws(K kernel, int cl, int smem_bytes, int groups, int threads, const RnnWsArgs &a,
                     cudaStream_t s, const char *what) {
    // Opt in to the kernel's dynamic shared memory once per (kernel, device) instead of on every launch
    // (a driver call).  Every instantiation has the same function type, so this template is
    // instantiated once and the kernels are told apart by their address.
    static const void *opted[8][96] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const void *key = reinterpret_cast<const void *>(kernel);
    bool known = false;
    int slot = -1;
    if (dev >= 0 && dev < 8) {
        for (int i = 0; i < 96; i++) {
            if (opted[dev][i] == key) { known = true; break; }
            if (opted[dev][i] == nullptr) { slot = i; break; }
        }
    }
    cudaError_t e = cudaSuccess;
    if (!known) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute(%d bytes): %s", what, smem_bytes, cudaGetErrorString(e));
            return TY_ECUDA;
        }
        if (slot >= 0) opted[dev][slot] = key;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(groups * cl);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (e != cudaSuccess) {
        set_error("%s: launch (cluster %d, %d threads, %d bytes): %s", what, cl, threads, smem_bytes,
                  cudaGetErrorString(e));
        return TY_ECUDA;
    }
    return check_launch(what);
}

template <int CELL, int H, int CL>
static int launch_rnn_ws_size(bool backward, const RnnWsArgs &a, cudaStream_t s) {
    const int groups = (a.N + kNB - 1) / kNB;
    using FL = FwdLayout<CELL, H, CL>;
    using BL = BwdLayout<CELL, H, CL>;
    return backward ? launch_ws(rnn_ws_backward_kernel<CELL, H, CL>, CL, BL::total, groups, BL::NCT + 32, a, s,
                                "rnn_ws_backward_kernel")
                    : launch_ws(rnn_ws_forward_kernel<CELL, H, CL>, CL, FL::total, groups, FL::NCT + 32, a, s,
                                "rnn_ws_forward_kernel");
}

// hidden sizes of the cluster kernels: multiples of 64 up to 448 (clusters of 8) and the odd
// multiples of 32 up to 224 (clusters of 4)
template <int CELL>
static int launch_rnn_ws(bool backward, const RnnWsArgs &a, int H, cudaStream_t s) {
#define TY_RNN_WS(HH, CLL) \
    case HH:               \
        return launch_rnn_ws_size<CELL, HH, CLL>(backward, a, s);
    switch (H) {
        TY_RNN_WS(64, 8) TY_RNN_WS(128, 8) TY_RNN_WS(192, 8) TY_RNN_WS(256, 8)
        TY_RNN_WS(320, 8) TY_RNN_WS(384, 8) TY_RNN_WS(448, 8)
        TY_RNN_WS(32, 4) TY_RNN_WS(96, 4) TY_RNN_WS(160, 4) TY_RNN_WS(224, 4)
        default:
            set_error("ty_rnn: hidden size %d has no cluster kernel (sizes: multiples of 64 up to 448; 32, 96, "
                      "160, 224)", H);
            return TY_EINVAL;
    }
#undef TY_RNN_WS
}

int launch_rnn_ws_lstm(bool backward, const RnnWsArgs &a, int H, cudaStream_t s);
int launch_rnn_ws_gru(bool backward, const RnnWsArgs &a, int H, cudaStream_t s);

}  // namespace ty
