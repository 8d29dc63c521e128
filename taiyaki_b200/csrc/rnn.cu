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
#include "rnn_common.cuh"

namespace ty {

// Forward.  H multiple of 64, H <= 256.  Threads: 32 * H/64 (warp w owns 8 units).
//
// The input projection of a step is needed in the gate math, on the critical
// path between two exchanges, and a global load issued one step ahead is not
// there in time (measured: 675 of 1620 cycles per step were this wait).  Each
// thread therefore streams ITS OWN 2*G values per step through a private slot
// of a shared-memory ring with cp.async, kXLook steps ahead: no registers, no
// scoreboard, no cross-thread synchronisation (a thread only reads what it
// copied itself, after cp.async.wait_group).
constexpr int kXRing = 4;    // ring slots (steps)
constexpr int kXLook = 3;    // steps between issue and use

template <int CELL, int H>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(H / 2, 1)
    rnn_forward_kernel(const RnnArgs a) {
    constexpr int G = Cell<CELL>::G;
    constexpr int U = H / kCluster;     // units per CTA
    constexpr int KT = H / 16;          // k tiles
    constexpr int NTHR = H / 2;

    // h_{t-1}, bf16, unit-major: 16 bytes (8 chunks) per unit; a warp's 8 units are 128
    // contiguous bytes, which is what makes the remote stores of the exchange cheap
    __shared__ __align__(128) __nv_bfloat16 hs[2][H][kNB];
    __shared__ __align__(16) float xs[kXRing][2][NTHR][4];   // [slot][column][thread][gate]
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / kCluster;
    const int unit = rank * U + warp * 8 + r;      // hidden unit of this thread's cells
    const int T = a.T, N = a.N;
    const int b0 = group * kNB + 2 * q;            // chunk of accumulator column 0
    const bool v0 = b0 < N, v1 = b0 + 1 < N;
    float *const gates_out = a.reserve;
    float *const cstate_out = a.reserve + (size_t)T * N * H * 4;

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

    for (int i = tid; i < 2 * kNB * H; i += NTHR) (&hs[0][0][0])[i] = __float2bfloat16(0.f);
#pragma unroll
    for (int sl = 0; sl < kXRing; sl++) {
        *reinterpret_cast<float4 *>(&xs[sl][0][tid][0]) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(&xs[sl][1][tid][0]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    cluster_sync_all();

    float cst[2] = {0.f, 0.f};    // LSTM cell state / GRU previous h (fp32)
    uint32_t phase = 0u;          // bit b = parity to wait for on full[b]
    // input bias of this thread's unit; it seeds the accumulators, so adding it is free
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias) {
#pragma unroll
        for (int g = 0; g < G; g++) bias[g] = a.bias[(size_t)g * H + unit];
    }

    auto tindex = [&](int s) { return a.reverse ? T - 1 - s : s; };
    // start the copies of step s's projection values into ring slot s % kXRing
    auto issue_x = [&](int s, int slot) {
        if (s < T) {
            const float *base = a.xproj + ((size_t)tindex(s) * N + b0) * (G * H) + unit;
            float *d0 = &xs[slot][0][tid][0], *d1 = &xs[slot][1][tid][0];
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (v0) cp_async4(d0 + g, base + (size_t)g * H);
                if (v1) cp_async4(d1 + g, base + (size_t)(G * H) + (size_t)g * H);
            }
        }
        cp_async_commit();     // one group per step, empty past the end: uniform counting
    };

    const uint32_t hs_base = smem_u32(&hs[0][0][0]);
    const uint32_t bar_base = smem_u32(&full[0]);
    // ldmatrix (transposed) source row for this lane: matrix lane>>3 = units 8*(lane>>3)..+7
    // of a 32-unit group, row lane&7 = one unit (16 bytes: its 8 chunks)
    const uint32_t ld_off = (uint32_t)(lane * 16);
    // destination of this thread's h pair (one unit, chunks 2q and 2q+1)
    const uint32_t send_off = (uint32_t)((unit * kNB + 2 * q) * 2);

    // SLOT = s % kXRing as a compile-time constant (ring slot, buffer parity)
    auto step = [&](const int s, auto slot_c) {
        constexpr int SLOT = decltype(slot_c)::value;
        const int t = tindex(s);
        constexpr int cur = SLOT & 1, nxt = cur ^ 1;
        issue_x(s + kXLook, (SLOT + kXLook) % kXRing);
        if (tid == 0 && s + 1 < T)      // arm the barrier that collects h_t (phase of step s+1)
            mbar_arrive_expect_tx(&full[nxt], kCluster * kNB * U * 2);
        cp_async_wait<kXLook>();        // this step's values have landed (issued kXLook steps ago)
        const float4 x0 = *reinterpret_cast<const float4 *>(&xs[SLOT][0][tid][0]);
        const float4 x1 = *reinterpret_cast<const float4 *>(&xs[SLOT][1][tid][0]);
        const float xp[4][2] = {{x0.x, x1.x}, {x0.y, x1.y}, {x0.z, x1.z}, {x0.w, x1.w}};
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
        // element e of tile m: gate 2m + (e >> 1), column e & 1.  Chain 0 is seeded with
        // the bias, chain 1 with the projection, so neither costs an addition.  (GRU: the
        // x side of the n gate must stay outside W_hn h; it is added below.)
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

        const uint32_t hcur = hs_base + (uint32_t)(cur * kNB * H * 2) + ld_off;
#pragma unroll
        for (int kp = 0; kp < KT / 2; kp++) {
            uint32_t bf[4];
            ldmatrix_x4_trans(bf, hcur + kp * 512);
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
            const size_t cell = ((size_t)t * N + b0 + col) * H + unit;
            float4 sv;
            if (CELL == kLstm) {
                const float gi = sigmoidf_(pre[0][col]);
                const float gf = sigmoidf_(pre[0][2 + col]);
                const float gg = tanhf_(pre[1][col]);
                const float go = sigmoidf_(pre[1][2 + col]);
                const float c = gf * cst[col] + gi * gg;
                cst[col] = c;
                hnew[col] = go * tanhf_(c);
                sv = make_float4(gi, gf, gg, go);
                if (valid) __stcs(cstate_out + cell, c);
            } else {
                const float hn = pre[1][col];
                const float gr = sigmoidf_(pre[0][col]);
                const float gz = sigmoidf_(pre[0][2 + col]);
                const float gn = tanhf_(xp[2][col] + bias[2] + gr * hn);
                const float h = (1.0f - gz) * gn + gz * cst[col];
                cst[col] = h;
                hnew[col] = h;
                sv = make_float4(gr, gz, gn, hn);
            }
            if (valid) {
                __stcs(reinterpret_cast<float4 *>(gates_out) + cell, sv);
                a.y[cell] = hnew[col];
            }
        }
        if (s + 1 < T) {
            // Exchange h_t: the thread's two cells are one unit of two adjacent chunks,
            // i.e. one 4-byte bf16 pair of the unit-major tile; it is stored straight from
            // registers into all 8 CTAs of the cluster (its own included) and the
            // receivers' mbarriers count the bytes.  A warp's 32 stores to one peer cover
            // 128 contiguous bytes (tools/dsmem_bench.cu: 460 cycles per exchange round
            // against 655 for eight 16-byte rows).
            const uint32_t pair = pack_bf16(hnew[0], hnew[1]);
            const uint32_t dst = hs_base + (uint32_t)(nxt * kNB * H * 2) + send_off;
            const uint32_t bar = bar_base + (uint32_t)(nxt * 8);
#pragma unroll
            for (uint32_t peer = 0; peer < kCluster; peer++)
                st_async_b32(mapa(dst, peer), pair, mapa(bar, peer));
        }
        if (a.y16) {     // bf16 copy of y: operand of the next layer's GEMMs
            if (v0) a.y16[((size_t)t * N + b0) * H + unit] = __float2bfloat16(hnew[0]);
            if (v1) a.y16[((size_t)t * N + b0 + 1) * H + unit] = __float2bfloat16(hnew[1]);
        }
    };

#pragma unroll
    for (int s0 = 0; s0 < kXLook; s0++) issue_x(s0, s0);
    static_assert(kXRing == 4 && kXLook < kXRing, "the loop below is unrolled by the ring size");
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
    cp_async_wait<0>();
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Forward, one-cell-per-thread variant.  Cluster of CL CTAs (8, or 16 with the
// non-portable cluster size); CTA j owns U = H/CL units; warp w owns ONE m16
// tile = 4 units x 4 gates, so the CTA has U/4 warps (8 at H=256, CL=8: two per
// SM sub-partition, each with half the instruction stream of the kernel above).
// Row g (= lane/4) of the tile is gate 2(g&1) of unit g/2 and row g+8 is gate
// 2(g&1)+1 of the same unit: a thread's accumulators hold two gates of one unit
// for two chunks; lanes l and l^4 swap two values so that each owns all four
// gates of ONE (unit, chunk) cell.  The projection x and the bias seed the two
// accumulator chains, so no addition follows the HMMAs.  GRU: row "gate 3" has
// zero weights and is seeded with x_n + b_n -- it carries the x-side of the n
// gate through the same swap.
template <int CELL, int H, int CL>
__global__ void __launch_bounds__(H / CL * 8, 1) rnn_forward_kernel2(const RnnArgs a) {
    constexpr int G = Cell<CELL>::G;
    constexpr int U = H / CL;           // units per CTA
    constexpr int KT = H / 16;          // k tiles
    constexpr int NTHR = U * 8;

    __shared__ __align__(128) __nv_bfloat16 hs[2][H][kNB];   // unit-major, see rnn_forward_kernel
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, q = lane & 3;
    const int odd = g8 & 1;                              // which gate pair this thread accumulates
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / CL;
    const int unit = rank * U + warp * 4 + (g8 >> 1);    // hidden unit of this thread's cell
    const int T = a.T, N = a.N;
    const int b0 = group * kNB + 2 * q;                  // chunk of accumulator column 0
    const int bc = b0 + odd;                             // chunk of this thread's cell
    const bool v0 = b0 < N, v1 = b0 + 1 < N, vc = bc < N;
    float *const gates_out = a.reserve;
    float *const cstate_out = a.reserve + (size_t)T * N * H * 4;

    // gate of accumulator slot 0 / 1 (rows g8, g8+8); -1 = no weights / no x
    const int gate_lo = 2 * odd, gate_hi = 2 * odd + 1;
    const bool w_hi = gate_hi < G;
    // x seeds: LSTM every slot; GRU: r, z in the even lanes; in the odd lanes slot 0
    // (W_hn h) stays pure and slot 1 (zero weights) carries x_n + b_n
    const int xg_lo = (CELL == kGru && odd) ? -1 : gate_lo;
    const int xg_hi = (CELL == kGru && odd) ? 2 : gate_hi;

    uint32_t A[KT][4];
    {
        const float *wlo = a.w_hh + ((size_t)gate_lo * H + unit) * H;
        const float *whi = a.w_hh + ((size_t)(w_hi ? gate_hi : 0) * H + unit) * H;
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
            const int k = 16 * kt + 2 * q;
            A[kt][0] = pack_bf16(wlo[k], wlo[k + 1]);
            A[kt][2] = pack_bf16(wlo[k + 8], wlo[k + 9]);
            A[kt][1] = w_hi ? pack_bf16(whi[k], whi[k + 1]) : 0u;
            A[kt][3] = w_hi ? pack_bf16(whi[k + 8], whi[k + 9]) : 0u;
        }
    }
    for (int i = tid; i < 2 * kNB * H; i += NTHR) (&hs[0][0][0])[i] = __float2bfloat16(0.f);
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    cluster_sync_all();

    float cst = 0.f;              // LSTM cell state / GRU previous h (fp32) of this thread's cell
    uint32_t phase = 0u;
    float bias_lo = 0.f, bias_hi = 0.f;
    if (a.bias) {
        if (xg_lo >= 0) bias_lo = a.bias[(size_t)xg_lo * H + unit];
        bias_hi = a.bias[(size_t)xg_hi * H + unit];
    }

    auto tindex = [&](int s) { return a.reverse ? T - 1 - s : s; };
    auto xaddr = [&](int s) { return a.xproj + ((size_t)tindex(s) * N + b0) * (G * H) + unit; };
    // x[slot][col]; see rnn_forward_kernel for `late`
    auto load_x = [&](float (&dst)[2][2], int s, uint32_t late) {
        if (s < T) {
            const float *base = xaddr(s) + (late & a.zero);
            if (xg_lo >= 0) {
                dst[0][0] = v0 ? ld_nc_pinned(base + (size_t)xg_lo * H) : 0.f;
                dst[0][1] = v1 ? ld_nc_pinned(base + (size_t)(G * H) + (size_t)xg_lo * H) : 0.f;
            }
            dst[1][0] = v0 ? ld_nc_pinned(base + (size_t)xg_hi * H) : 0.f;
            dst[1][1] = v1 ? ld_nc_pinned(base + (size_t)(G * H) + (size_t)xg_hi * H) : 0.f;
        }
    };
    auto prefetch_x = [&](int s) {
        if (s < T) {
            const float *base = xaddr(s);
            if (xg_lo >= 0) {
                if (v0) prefetch_l2(base + (size_t)xg_lo * H);
                if (v1) prefetch_l2(base + (size_t)(G * H) + (size_t)xg_lo * H);
            }
            if (v0) prefetch_l2(base + (size_t)xg_hi * H);
            if (v1) prefetch_l2(base + (size_t)(G * H) + (size_t)xg_hi * H);
        }
    };

    const uint32_t hs_base = smem_u32(&hs[0][0][0]);
    const uint32_t bar_base = smem_u32(&full[0]);
    const uint32_t ld_off = (uint32_t)(lane * 16);
    // h exchange: lanes l, l^4 hold chunks 2q, 2q+1 of one unit -> one bf16 pair of the
    // unit-major tile; the even lane serves the lower half of the peers, the odd one the
    // upper half.  A warp's stores to one peer cover its private 64 contiguous bytes.
    const uint32_t send_off = (uint32_t)((unit * kNB + 2 * q) * 2);
    const uint32_t peer0 = odd ? (uint32_t)(CL / 2) : 0u;

    auto step = [&](const int s, float (&xp)[2][2], float (&xn)[2][2]) {
        const int t = tindex(s);
        const int cur = s & 1, nxt = cur ^ 1;
        prefetch_x(s + 3);
        if (tid == 0 && s + 1 < T) mbar_arrive_expect_tx(&full[nxt], CL * kNB * U * 2);
        if (s > 0) {
            mbar_wait(&full[cur], (phase >> cur) & 1u);
            phase ^= 1u << cur;
        }
        // c-fragment element e: slot e >> 1 (row g8 / g8+8), column e & 1
        float acc[2][4];
        acc[0][0] = (CELL == kGru && odd) ? 0.f : xp[0][0];
        acc[0][1] = (CELL == kGru && odd) ? 0.f : xp[0][1];
        acc[0][2] = xp[1][0]; acc[0][3] = xp[1][1];
        acc[1][0] = bias_lo; acc[1][1] = bias_lo;
        acc[1][2] = bias_hi; acc[1][3] = bias_hi;

        const uint32_t hcur = hs_base + (uint32_t)(cur * kNB * H * 2) + ld_off;
#pragma unroll
        for (int kp = 0; kp < KT / 2; kp++) {
            uint32_t bf[4];
            ldmatrix_x4_trans(bf, hcur + kp * 512);
            mma_bf16(acc[0], A[2 * kp], bf[0], bf[1]);
            mma_bf16(acc[1], A[2 * kp + 1], bf[2], bf[3]);
        }
        float pre[4];
#pragma unroll
        for (int e = 0; e < 4; e++) pre[e] = acc[0][e] + acc[1][e];
        // swap: even lanes keep column 0 and give column 1, odd lanes the reverse
        const float give0 = odd ? pre[0] : pre[1];
        const float give1 = odd ? pre[2] : pre[3];
        const float got0 = __shfl_xor_sync(kFullMask, give0, 4);
        const float got1 = __shfl_xor_sync(kFullMask, give1, 4);
        // v[0..3] = pre-activations of gates 0..3 of this thread's cell
        const float p0 = odd ? got0 : pre[0];
        const float p1 = odd ? got1 : pre[2];
        const float p2 = odd ? pre[1] : got0;
        const float p3 = odd ? pre[3] : got1;

        const size_t cell = ((size_t)t * N + bc) * H + unit;
        float hnew;
        float4 sv;
        if (CELL == kLstm) {
            const float gi = sigmoidf_(p0);
            const float gf = sigmoidf_(p1);
            const float gg = tanhf_(p2);
            const float go = sigmoidf_(p3);
            const float c = gf * cst + gi * gg;
            cst = c;
            hnew = go * tanhf_(c);
            sv = make_float4(gi, gf, gg, go);
            if (vc) __stcs(cstate_out + cell, c);
        } else {
            const float gr = sigmoidf_(p0);
            const float gz = sigmoidf_(p1);
            const float gn = tanhf_(p3 + gr * p2);
            const float h = (1.0f - gz) * gn + gz * cst;
            cst = h;
            hnew = h;
            sv = make_float4(gr, gz, gn, p2);
        }
        if (vc) {
            __stcs(reinterpret_cast<float4 *>(gates_out) + cell, sv);
            a.y[cell] = hnew;
        }
        const uint32_t late_tok = __float_as_uint(hnew);
        if (s + 1 < T) {
            const float other = __shfl_xor_sync(kFullMask, hnew, 4);
            const uint32_t pair = odd ? pack_bf16(other, hnew) : pack_bf16(hnew, other);
            const uint32_t dst = hs_base + (uint32_t)(nxt * kNB * H * 2) + send_off;
            const uint32_t bar = bar_base + (uint32_t)(nxt * 8);
#pragma unroll
            for (uint32_t p = 0; p < CL / 2; p++)
                st_async_b32(mapa(dst, peer0 + p), pair, mapa(bar, peer0 + p));
        }
        if (a.y16 && vc) a.y16[cell] = __float2bfloat16(hnew);
        load_x(xn, s + 1, late_tok);
    };

    float xa[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, xb[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    load_x(xa, 0, 0u);
    prefetch_x(1);
    prefetch_x(2);
    int s = 0;
    for (; s + 1 < T; s += 2) {
        step(s, xa, xb);
        step(s + 1, xb, xa);
    }
    if (s < T) step(s, xa, xb);
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Backward (BPTT).  Same ownership as forward.  K dimension of the per-step
// product is this CTA's G*U gate rows (padded to a multiple of 16); M is all H
// hidden units; the partial results are reduce-scattered to the owners.
template <int CELL, int H>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(H / 2, 1)
    rnn_backward_kernel(const RnnArgs a) {
    constexpr int G = Cell<CELL>::G;
    constexpr int U = H / kCluster;
    constexpr int KL = (G * U + 15) / 16 * 16;   // local gate rows, padded
    constexpr int KT = KL / 16;
    constexpr int DS = KL + 8;                   // padded row (bf16)
    constexpr int NTHR = H / 2;
    constexpr int MT = 4;                        // m tiles (16 units) per warp: (H/16)/(H/64)

    __shared__ __align__(16) __nv_bfloat16 ds[kNB][DS];
    __shared__ __align__(16) float rs[2][kCluster][U][kNB];
    __shared__ __align__(16) float4 gs[kXRing][2][NTHR];     // saved gates [slot][column][thread]
    __shared__ __align__(16) float vs[kXRing][2][2][NTHR];   // dy | c_t (LSTM) or h_{t-1} (GRU)
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
    const float4 *const gates_in = reinterpret_cast<const float4 *>(a.reserve);
    const float *const cstate_in = a.reserve + (size_t)T * N * H * 4;

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
#pragma unroll
    for (int sl = 0; sl < kXRing; sl++) {
#pragma unroll
        for (int col = 0; col < 2; col++) {
            gs[sl][col][tid] = make_float4(0.f, 0.f, 0.f, 0.f);
            vs[sl][0][col][tid] = 0.f;
            vs[sl][1][col][tid] = 0.f;
        }
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    cluster_sync_all();

    uint32_t phase = 0u;
    float carry[2] = {0.f, 0.f};   // LSTM: dL/dc carried back; GRU: z * dL/dh carried back
    float dbacc[G];                // bias gradient of this thread's unit: sum over time and its chunks
#pragma unroll
    for (int g = 0; g < G; g++) dbacc[g] = 0.f;
    auto tindex = [&](int sf) { return a.reverse ? T - 1 - sf : sf; };   // forward step -> time

    // Per-step inputs (saved gates, c_t or h_{t-1}, incoming gradient) stream through
    // thread-private slots of a shared-memory ring with cp.async, kXLook steps ahead
    // (see rnn_forward_kernel).  LSTM: c_{t-1} is simply the next slot's c.
    auto issue_in = [&](int s, int slot) {
        if (s < T) {
            const int sf = T - 1 - s;                 // forward step being differentiated
            const int t = tindex(sf);
#pragma unroll
            for (int col = 0; col < 2; col++) {
                if (col == 0 ? v0 : v1) {
                    const size_t cell = ((size_t)t * N + b0 + col) * H + unit;
                    cp_async16(&gs[slot][col][tid], gates_in + cell);
                    cp_async4(&vs[slot][0][col][tid], a.dy + cell);
                    if (CELL == kLstm) {
                        cp_async4(&vs[slot][1][col][tid], cstate_in + cell);
                    } else if (sf > 0) {
                        const size_t pcell = ((size_t)tindex(sf - 1) * N + b0 + col) * H + unit;
                        cp_async4(&vs[slot][1][col][tid], a.y + pcell);     // h_{t-1}
                    }
                }
            }
        }
        cp_async_commit();
    };

    const uint32_t ds_base = smem_u32(&ds[0][0]);
    const uint32_t ld_off = (uint32_t)(((lane & 7) * DS + 8 * (lane >> 3)) * 2);
    const uint32_t rs_base = smem_u32(&rs[0][0][0][0]);
    const uint32_t bar_base = smem_u32(&full[0]);

    auto step = [&](const int s, auto slot_c) {
        constexpr int SLOT = decltype(slot_c)::value;
        constexpr int cur = SLOT & 1, nxt = cur ^ 1;
        const int sf = T - 1 - s;
        const int t = tindex(sf);
        issue_in(s + kXLook, (SLOT + kXLook) % kXRing);
        if (tid == 0 && s + 1 < T) mbar_arrive_expect_tx(&full[nxt], kCluster * U * kNB * 4);
        // this step's group and the next one's (for c_{t-1}) have landed
        cp_async_wait<kXLook - 1>();
        struct { float4 gt[2]; float c[2]; float dy[2]; float prev[2]; } in;
#pragma unroll
        for (int col = 0; col < 2; col++) {
            in.gt[col] = gs[SLOT][col][tid];
            in.dy[col] = vs[SLOT][0][col][tid];
            if (CELL == kLstm) {
                in.c[col] = vs[SLOT][1][col][tid];
                in.prev[col] = sf > 0 ? vs[(SLOT + 1) % kXRing][1][col][tid] : 0.f;   // c_{t-1}
            } else {
                in.c[col] = 0.f;
                in.prev[col] = sf > 0 ? vs[SLOT][1][col][tid] : 0.f;                  // h_{t-1}
            }
        }

        float dh[2] = {in.dy[0], in.dy[1]};
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

#pragma unroll
        for (int col = 0; col < 2; col++) {
            const bool valid = col == 0 ? v0 : v1;
            const size_t cell = ((size_t)t * N + b0 + col) * H + unit;
            const size_t xrow = ((size_t)t * N + b0 + col) * (G * H) + unit;
            float dg[G];
            if (CELL == kLstm) {
                const float gi = in.gt[col].x, gf = in.gt[col].y, gg = in.gt[col].z,
                            go = in.gt[col].w;
                const float tc = tanhf_(in.c[col]);
                const float d = dh[col];
                const float dc = carry[col] + d * go * (1.0f - tc * tc);
                dg[3] = d * tc * go * (1.0f - go);
                dg[0] = dc * gg * gi * (1.0f - gi);
                dg[2] = dc * gi * (1.0f - gg * gg);
                dg[1] = dc * in.prev[col] * gf * (1.0f - gf);
                carry[col] = dc * gf;
#pragma unroll
                for (int g = 0; g < G; g++)
                    ds[2 * q + col][g * U + ul] = __float2bfloat16(dg[g]);
            } else {
                const float gr = in.gt[col].x, gz = in.gt[col].y, gn = in.gt[col].z,
                            hn = in.gt[col].w;
                const float d = dh[col] + carry[col];
                const float dn = d * (1.0f - gz) * (1.0f - gn * gn);      // d n_pre
                dg[1] = d * (in.prev[col] - gn) * gz * (1.0f - gz);
                dg[0] = dn * hn * gr * (1.0f - gr);
                dg[2] = dn;                                               // x-side n gradient
                carry[col] = d * gz;
                const float dhn = dn * gr;                                // hidden-side n gradient
                if (valid) {
                    if (a.dhn16) a.dhn16[cell] = __float2bfloat16(dhn);
                    else __stcs(a.dhn + cell, dhn);
                }
                // the recurrent product uses the hidden-side gradient for gate n
                ds[2 * q + col][0 * U + ul] = __float2bfloat16(dg[0]);
                ds[2 * q + col][1 * U + ul] = __float2bfloat16(dg[1]);
                ds[2 * q + col][2 * U + ul] = __float2bfloat16(dhn);
            }
            if (valid) {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (a.dxproj16) a.dxproj16[xrow + (size_t)g * H] = __float2bfloat16(dg[g]);
                    else __stcs(a.dxproj + xrow + (size_t)g * H, dg[g]);
                    dbacc[g] += dg[g];
                }
            }
        }
        if (s + 1 < T) {
            __syncthreads();
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
            const uint32_t rs_nxt = rs_base + (uint32_t)(((nxt * kCluster + rank) * U * kNB) * 4);
            const uint32_t bar = bar_base + (uint32_t)(nxt * 8);
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
                    st_async_v2(mapa(local, dest), x, y, mapa(bar, dest));
                }
            }
            // ds is rewritten by the next step's gate gradients: every warp of
            // this CTA must be past its ldmatrix reads first.  The next wait on
            // `full` implies it (the barrier needs this CTA's own partials too),
            // except for the pad rows, which are never rewritten.
        }
    };

    for (int s0 = 0; s0 < kXLook; s0++) issue_in(s0, s0);
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
    cp_async_wait<0>();
    if (a.dbias) {
#pragma unroll
        for (int g = 0; g < G; g++) atomicAdd(a.dbias + (size_t)g * H + unit, dbacc[g]);
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------
// Forward variant: 0 = two cells per thread (rnn_forward_kernel), 1 = one cell per
// thread, cluster of 8, 2 = one cell per thread, cluster of 16 where it fits.
static int fwd_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TY_RNN_FWD");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <typename K>
static cudaError_t launch_cluster(K kernel, int grid, int block, int cl, cudaStream_t s,
                                  const RnnArgs &a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

// Can `groups` clusters of 16 CTAs be co-resident?  (One per GPC at most.)
template <typename K>
static bool cluster16_fits(K kernel, int block, int groups) {
    static int max_clusters = -1;
    if (max_clusters < 0) {
        max_clusters = 0;
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) ==
            cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(16 * 64);
            cfg.blockDim = dim3(block);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 16;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess) max_clusters = n;
        }
        cudaGetLastError();
        if (getenv("TY_RNN_DEBUG")) fprintf(stderr, "ty_rnn: max active clusters of 16: %d\n", max_clusters);
    }
    return groups <= max_clusters;
}

template <int CELL>
static int launch_rnn(bool backward, const RnnArgs &a, int H, cudaStream_t s) {
    const int groups = (a.N + kNB - 1) / kNB;
    const dim3 grid(groups * kCluster);
    const int variant = backward ? 0 : fwd_variant();
    if (variant >= 1 && H == 256) {
        cudaError_t e;
        if (variant == 2 && cluster16_fits(rnn_forward_kernel2<CELL, 256, 16>, 128, groups)) {
            e = launch_cluster(rnn_forward_kernel2<CELL, 256, 16>, groups * 16, 128, 16, s, a);
        } else {
            e = launch_cluster(rnn_forward_kernel2<CELL, 256, 8>, groups * 8, 256, 8, s, a);
        }
        if (e != cudaSuccess) {
            set_error("rnn_forward_kernel2 launch: %s", cudaGetErrorString(e));
            return TY_ECUDA;
        }
        return TY_OK;
    }
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
    (void)cell;
    return (size_t)T * N * 5 * H * sizeof(float);
}

// bf16 side outputs: see include/taiyaki_b200.h (ty_rnn_forward_ex / ty_rnn_backward_ex)
extern "C" int ty_rnn_forward_ex(int cell, const float *xproj, const float *bias,
                                 const float *w_hh, int T, int N, int H, int reverse, float *y,
                                 void *y_bf16, void *reserve, void *stream) {
    if (int rc = check_shape(T, N, H, xproj, w_hh, y)) return rc;
    if (!reserve) { set_error("ty_rnn_forward_ex: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.xproj = xproj; a.bias = bias; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse; a.y = y;
    a.y16 = static_cast<__nv_bfloat16 *>(y_bf16);
    a.reserve = static_cast<float *>(reserve);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return cell == kLstm ? launch_rnn<kLstm>(false, a, H, s) : launch_rnn<kGru>(false, a, H, s);
}

extern "C" int ty_rnn_backward_ex(int cell, const float *dy, const float *w_hh, int T, int N,
                                  int H, int reverse, const float *y, const void *reserve,
                                  void *dxproj, void *dhn, int grads_bf16, float *dbias,
                                  void *stream) {
    if (int rc = check_shape(T, N, H, dy, w_hh, dxproj)) return rc;
    if (!reserve || (cell == kGru && (!y || !dhn))) {
        set_error("ty_rnn_backward_ex: null pointer");
        return TY_EINVAL;
    }
    RnnArgs a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.reserve = const_cast<float *>(static_cast<const float *>(reserve));
    if (grads_bf16) {
        a.dxproj16 = static_cast<__nv_bfloat16 *>(dxproj);
        a.dhn16 = static_cast<__nv_bfloat16 *>(dhn);
    } else {
        a.dxproj = static_cast<float *>(dxproj);
        a.dhn = static_cast<float *>(dhn);
    }
    a.dbias = dbias;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return cell == kLstm ? launch_rnn<kLstm>(true, a, H, s) : launch_rnn<kGru>(true, a, H, s);
}

extern "C" int ty_lstm_forward(const float *xproj, const float *bias, const float *w_hh, int T,
                               int N, int H, int reverse, float *y, void *reserve,
                               void *stream) {
    if (int rc = check_shape(T, N, H, xproj, w_hh, y)) return rc;
    if (!reserve) { set_error("ty_lstm_forward: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.xproj = xproj; a.bias = bias; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse; a.y = y;
    a.reserve = static_cast<float *>(reserve);
    return launch_rnn<kLstm>(false, a, H, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_lstm_backward(const float *dy, const float *w_hh, int T, int N, int H,
                                int reverse, const float *y, const void *reserve, float *dxproj,
                                float *dbias, void *stream) {
    if (int rc = check_shape(T, N, H, dy, w_hh, dxproj)) return rc;
    if (!reserve) { set_error("ty_lstm_backward: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.reserve = const_cast<float *>(static_cast<const float *>(reserve));
    a.dxproj = dxproj; a.dbias = dbias;
    return launch_rnn<kLstm>(true, a, H, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_gru_forward(const float *xproj, const float *bias, const float *w_hh, int T,
                              int N, int H, int reverse, float *y, void *reserve,
                              void *stream) {
    if (int rc = check_shape(T, N, H, xproj, w_hh, y)) return rc;
    if (!reserve) { set_error("ty_gru_forward: reserve is null"); return TY_EINVAL; }
    RnnArgs a{};
    a.xproj = xproj; a.bias = bias; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse; a.y = y;
    a.reserve = static_cast<float *>(reserve);
    return launch_rnn<kGru>(false, a, H, static_cast<cudaStream_t>(stream));
}

extern "C" int ty_gru_backward(const float *dy, const float *w_hh, int T, int N, int H,
                               int reverse, const float *y, const void *reserve, float *dxproj,
                               float *dhn, float *dbias, void *stream) {
    if (int rc = check_shape(T, N, H, dy, w_hh, dxproj)) return rc;
    if (!reserve || !y || !dhn) { set_error("ty_gru_backward: null pointer"); return TY_EINVAL; }
    RnnArgs a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.reserve = const_cast<float *>(static_cast<const float *>(reserve));
    a.dxproj = dxproj; a.dhn = dhn; a.dbias = dbias;
    return launch_rnn<kGru>(true, a, H, static_cast<cudaStream_t>(stream));
}
