// crf_flipflop.cu -- label-constrained flip-flop CRF forward/backward/posterior
// on sm_100a.  Replaces taiyaki/ctc/c_crf_flipflop.c:43-516 and
// c_cat_mod_flipflop.c:37-582 (the MOD template flag adds the per-move
// modified-base term).
//
// Structure (DESIGN.md "CRF kernels"):
//   crf_chain_kernel   one CTA per (chunk, direction).  The sequence positions of the chunk are
//                      striped over the threads, P consecutive positions per thread in
//                      registers; time is the only sequential dimension and the kernel is bound
//                      by the per-step dependency latency, so the step is written for minimum
//                      instruction count: the DP runs in the log2 domain (bare ex2/lg2 SFU ops);
//                      a "transformer" warp turns the raw score row (cp.async ring) into
//                      w*sharp*log2(e) - c one step ahead, so a position costs 2 LDS gathers and
//                      8 arithmetic instructions; masked positions point at a -1e30 pad slot
//                      instead of being selected away; the normaliser c is the block max of the
//                      vector two steps earlier (any finite per-(t,chunk) scalar is valid: the
//                      shifts are summed into the score, c_crf_flipflop.c:73-77, :124) so no
//                      reduction sits on the dependency chain.  alpha_t / beta_{t+1} rows and
//                      their accumulated offsets go to the HBM workspace.
//   crf_post_kernel    posterior of c_crf_flipflop.c:372-413 / c_cat_mod_flipflop.c:419-468 from
//                      the spilled rows, one warp per (block, chunk) row, position-major.
// Since round 2 this pair serves cost-only calls and chunks beyond the range of the fused
// kernel (crf_fused.cu), which runs the same DP step and the same posterior scatter in one launch.
#include "crf_common.cuh"

namespace ty {

// One chain.  Warp-specialised: warps 0..ncw-1 run the DP; the LAST warp is the
// "transformer": it owns the cp.async ring, turns raw score rows into
// w*sharp*log2(e) - c one step ahead, and keeps the offset bookkeeping, so the
// DP warps execute nothing but the recurrence.
template <int P, bool MOD, int DIR>
__device__ __forceinline__ void crf_chain_body(const CrfArgs &a, const int b, const int L,
                                               const int off) {
    __shared__ __align__(16) float raw[kRing][kRowPad];
    __shared__ __align__(16) float tr[2][kRowPad];
    __shared__ __align__(16) float bnd[2][32];
    __shared__ __align__(16) float wmaxs[2][32];
    __shared__ float s_end;

    const int tid = (int)pinned_tid(), lane = tid & 31, warp = tid >> 5;
    const int ncw = (blockDim.x >> 5) - 1;          // DP warps
    const bool is_tx = warp == ncw;
    const int S = a.ntrans;
    const int nblk = a.nblk;
    const size_t ld = (size_t)a.nbatch * S;
    const int p0 = tid * P;
    // shared-window addresses used inside the step loop (see lds_v_f32)
    const unsigned tr_u32 = opaque((unsigned)__cvta_generic_to_shared(&tr[0][0]));
    const unsigned raw_u32 = opaque((unsigned)__cvta_generic_to_shared(&raw[0][0]) + lane * 4);
    const unsigned bnd_u32 = opaque((unsigned)__cvta_generic_to_shared(&bnd[0][0]));
    const unsigned wmaxs_u32 = opaque((unsigned)__cvta_generic_to_shared(&wmaxs[0][0]));

    // ---- DP state: byte offsets into a transformed row (pad slot when invalid) ----
    int st[P], mv[P], mm[P];
    float mf[P];
    float al[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        const int p = p0 + i;
        st[i] = kPadSlot * 4; mv[i] = kPadSlot * 4; mm[i] = 0; mf[i] = 0.f;
        al[i] = kNegLarge;
        if (!is_tx) {
            if (p < L) st[i] = a.stayidx[off + p] * 4;
            // DIR 0 (forward): move INTO p from p-1.  DIR 1: move OUT of p to p+1.
            const int q = DIR == 0 ? p - 1 : p;
            if (p < L && q >= 0 && q < L - 1) {
                mv[i] = a.moveidx[off - b + q] * 4;
                if (MOD) {
                    mm[i] = a.modmoveidx[off - b + q] * 4;
                    mf[i] = a.modmovefact[off - b + q];
                }
            }
            // c_crf_flipflop.c:113-116 / :220-224 point priors
            if (p == (DIR == 0 ? 0 : L - 1)) al[i] = 0.f;
        }
    }

    // ---- transformer state ----
    const bool l0 = lane < S, l1 = lane + 32 < S;
    const float sc0 = (lane < a.nsharp ? a.sharp : 1.0f) * kLog2e;
    const float sc1 = (lane + 32 < a.nsharp ? a.sharp : 1.0f) * kLog2e;
    const bool can0 = lane < a.ncan, can1 = lane + 32 < a.ncan;
    const long long tstep = DIR == 0 ? (long long)ld : -(long long)ld;
    // raw row k lives at src(k) = lp + t(k)*ld ; t(k) = k or nblk-1-k
    const float *src = a.logprob + (size_t)b * S + (size_t)(DIR == 0 ? 0 : nblk - 1) * ld + lane;
    auto issue_row = [&](int k) {       // src points at row k
        if (k < nblk) {
            if (l0) cp_async4(&raw[k & (kRing - 1)][lane], src);
            if (l1) cp_async4(&raw[k & (kRing - 1)][lane + 32], src + 32);
        }
        cp_async_commit();
        src += tstep;
    };
    auto transform_row = [&](int k, int par, float c) {
        const unsigned ra = raw_u32 + (unsigned)(k & (kRing - 1)) * (kRowPad * 4);
        const unsigned ta = tr_u32 + (unsigned)par * (kRowPad * 4) + lane * 4;
        if (l0) {
            const float w = lds_v_f32(ra);
            sts_v_f32(ta, can0 ? fmaf(w, sc0, -c) : w * sc0);
        }
        if (l1) {
            const float w = lds_v_f32(ra + 128);
            sts_v_f32(ta + 128, can1 ? fmaf(w, sc1, -c) : w * sc1);
        }
    };
    float coff_run = 0.f;             // accumulated offset of the vector in registers
    float c_cur = 0.f, c_prev = 0.f;  // shifts inside the current / previous row
    float part = 0.f;                 // lane-striped partial sums of the shifts
    float *coff = nullptr;
    if (is_tx) {
#pragma unroll
        for (int k = 0; k < kDepth; k++) issue_row(k);
        wmaxs[0][lane] = lane < ncw ? 0.f : -3.0e38f;
        wmaxs[1][lane] = lane < ncw ? 0.f : -3.0e38f;
        if (lane == 0) { tr[0][kPadSlot] = kNegLarge; tr[1][kPadSlot] = kNegLarge; }
        cp_async_wait<kDepth - 1>();      // own copies of row 0 have landed
        transform_row(0, 0, 0.f);
        if (a.want_grad)
            coff = a.coff + ((size_t)DIR * a.nbatch + b) * nblk + (DIR == 0 ? 0 : nblk - 1);
    } else {
        // boundary words of the initial vector, read by step 0
        if (DIR == 0) {
            if (lane == 31) bnd[1][warp] = al[P - 1];
        } else {
            if (lane == 0) bnd[1][warp] = al[0];
        }
    }
    float pend_max = 0.f;             // warp max of the vector in registers
    __syncthreads();

    float *dst = nullptr;             // spill row of the current step
    const long long dstep = DIR == 0 ? (long long)a.Ls : -(long long)a.Ls;
    if (a.want_grad && !is_tx)
        dst = (DIR == 0 ? a.fwd_ws : a.bwd_ws) +
              ((size_t)b * nblk + (DIR == 0 ? 0 : nblk - 1)) * a.Ls + p0;

    // One time step; PAR = k & 1 is a compile-time constant (loop unrolled by 2)
    // so the row / boundary buffers are addressed with immediates.
    auto step = [&](const int k, auto par_c) {
        constexpr int PAR = decltype(par_c)::value;
        if (is_tx) {
            issue_row(k + kDepth);
            // Shift of the next row.  The freshest block max visible without a
            // reduction on the chain is that of the vector two steps ago,
            // m_{k-1}; two shifts (rows k-1, k) were applied since, so the shift
            // that re-centres is m_{k-1} - c_{k-1} - c_k.  (Subtracting the stale
            // max alone is an unstable recurrence.)
            const float c_next = warp_max(lds_v_f32(wmaxs_u32 + (PAR ^ 1) * 128 + lane * 4)) - c_cur - c_prev;
            if (a.want_grad) {
                if (lane == 0) *coff = coff_run;
                coff += DIR == 0 ? 1 : -1;
            }
            if (lane == (k & 31)) part += c_cur;
            coff_run += c_cur;
            c_prev = c_cur;
            c_cur = c_next;
            cp_async_wait<kDepth - 1>();          // own copies of row k+1 have landed
            if (k + 1 < nblk) transform_row(k + 1, PAR ^ 1, c_next);
        } else {
            if (lane == 0) sts_v_f32(wmaxs_u32 + PAR * 128 + warp * 4, pend_max);
            // spill alpha_t (DIR 0) / beta_{t+1} (DIR 1) for the posterior kernel
            if (a.want_grad) {
                if (P % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < P; i += 4)
                        if (p0 + i < L)
                            __stcs(reinterpret_cast<float4 *>(dst + i),
                                   make_float4(al[i], al[i + 1], al[i + 2], al[i + 3]));
                } else {
#pragma unroll
                    for (int i = 0; i < P; i++)
                        if (p0 + i < L) __stcs(dst + i, al[i]);
                }
                dst += dstep;
            }
            // gathers from the transformed row first (volatile: they keep this order),
            // then the neighbour value across the thread boundary
            const unsigned row = tr_u32 + PAR * (kRowPad * 4);
            float gs[P], gm[P];
#pragma unroll
            for (int i = 0; i < P; i++) {
                gs[i] = lds_v_f32(row + st[i]);
                gm[i] = lds_v_f32(row + mv[i]);
            }
            if (MOD) {
#pragma unroll
                for (int i = 0; i < P; i++) gm[i] = fmaf(lds_v_f32(row + mm[i]), mf[i], gm[i]);
            }
            float nb;
            if (DIR == 0) {
                nb = __shfl_up_sync(kFullMask, al[P - 1], 1);
                if (lane == 0) nb = warp > 0 ? lds_v_f32(bnd_u32 + (PAR ^ 1) * 128 + (warp - 1) * 4) : kNegLarge;
            } else {
                nb = __shfl_down_sync(kFullMask, al[0], 1);
                if (lane == 31) nb = warp + 1 < ncw ? lds_v_f32(bnd_u32 + (PAR ^ 1) * 128 + (warp + 1) * 4) : kNegLarge;
            }
            float nw[P];
            float tmax = -3.0e38f;
#pragma unroll
            for (int i = 0; i < P; i++) {
                float other;
                if (DIR == 0) other = (i == 0) ? nb : al[i - 1];
                else other = (i == P - 1) ? nb : al[i + 1];
                nw[i] = logaddexp2(al[i] + gs[i], other + gm[i]);
                tmax = fmaxf(tmax, nw[i]);
            }
#pragma unroll
            for (int i = 0; i < P; i++) al[i] = nw[i];
            if (DIR == 0) {
                if (lane == 31) sts_v_f32(bnd_u32 + PAR * 128 + warp * 4, al[P - 1]);
            } else {
                if (lane == 0) sts_v_f32(bnd_u32 + PAR * 128 + warp * 4, al[0]);
            }
            pend_max = warp_max(tmax);
        }
        __syncthreads();
    };

    int k = 0;
    for (; k + 1 < nblk; k += 2) {
        step(k, std::integral_constant<int, 0>{});
        step(k + 1, std::integral_constant<int, 1>{});
    }
    if (k < nblk) step(k, std::integral_constant<int, 0>{});

    // c_crf_flipflop.c:131-132 / :234: final position (forward) or first (backward)
    const int pend = DIR == 0 ? L - 1 : 0;
    if (!is_tx && pend >= p0 && pend < p0 + P) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < P; i++)
            if (p0 + i == pend) v = al[i];
        s_end = v;
    }
    __syncthreads();
    if (is_tx) {
        double tot = (double)part;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
        if (lane == 0) {
            const float score2 = (float)(tot + (double)s_end);       // log2 units
            if (a.want_grad) a.fb[2 * b + DIR] = score2;
            else a.score_out[b] = a.score_scale * kLn2 * score2;
        }
    }
}

template <int P, bool MOD>
__global__ void __launch_bounds__(P <= 4 ? 1024 : 544) crf_chain_kernel(const CrfArgs a) {
    __shared__ int s_off;
    const int b = a.want_grad ? (blockIdx.x >> 1) : blockIdx.x;
    const int dir = a.want_grad ? (blockIdx.x & 1) : 0;

    // prefix sum of seqlen (c_crf_flipflop.c:447-451)
    if (threadIdx.x < 32) {
        int s = 0;
        for (int i = threadIdx.x; i < b; i += 32) s += a.seqlen[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
        if (threadIdx.x == 0) s_off = s;
    }
    __syncthreads();
    const int off = s_off;
    const int L = a.seqlen[b];
    if (threadIdx.x == 0 && dir == 0 && a.seqoff) a.seqoff[b] = off;
    if (L <= 0) {   // c_crf_flipflop.c:269-272, :458-464
        if (threadIdx.x == 0) {
            if (a.want_grad) a.fb[2 * b + dir] = 0.f;
            else a.score_out[b] = 0.f;
        }
        return;
    }
    if (dir == 0) crf_chain_body<P, MOD, 0>(a, b, L, off);
    else crf_chain_body<P, MOD, 1>(a, b, L, off);
}

// ---------------------------------------------------------------------------
// Posterior of the chain kernel's spill (c_crf_flipflop.c:372-413 / c_cat_mod_flipflop.c:419-468).
// grid = (row tiles, chunks); one WARP per row (block, chunk), position-major: lane l takes
// positions l, l + 32, ... -- alpha_t, beta_{t+1} and the transition indices are read where they
// lie, coalesced, with no staging and therefore no limit on the chunk length -- and adds
// 2^(alpha + beta + w - Z_t) of the position's stay and move edges into ITS OWN column of a
// [transition][lane] table in shared memory (plain read-modify-writes, no atomics, a fixed order);
// the 32 columns of a transition are then summed with a rotated start.  Z_t is the analytic
// normaliser (total score minus the accumulated offsets of the two rows); the canonical
// transitions are renormalised to sum to one exactly like the reference's softmax
// (c_crf_flipflop.c:401), so Z_t only has to be close.
// (Round 1 walked transition-sorted entry lists from shared-memory copies of the rows; beyond the
// rows that fit it gathered them from global memory, 4 bytes per 32-byte sector: 170 ms at
// nblk 8000 x 4400 positions, 0.9 GB/s.  The same scatter runs inside crf_fused.cu.)
constexpr int kPostWarps = 8;

template <bool MOD>
__global__ void __launch_bounds__(kPostWarps * 32) crf_post_kernel(const CrfArgs a) {
    extern __shared__ __align__(16) float cols_all[];      // [kPostWarps][S][32]
    __shared__ float q[kPostWarps][kRowPad];                // w2 - Z_t (canonical), w2 (mod)

    const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kPostWarps;
    const int S = a.ntrans, Ls = a.Ls;
    const int L = a.seqlen[b];
    const int t = t0 + r;

    if (L <= 0) {
        if (t < a.nblk)
            for (int s = lane; s < S; s += 32)
                a.grad_out[((size_t)t * a.nbatch + b) * S + s] = 0.f;
        if (t0 == 0 && tid == 0) a.score_out[b] = 0.f;
        return;
    }
    // total log2 score = mean of forward and backward (c_crf_flipflop.c:482-491)
    const float score2 = 0.5f * (a.fb[2 * b] + a.fb[2 * b + 1]);
    if (t0 == 0 && tid == 0) a.score_out[b] = a.score_scale * kLn2 * score2;
    if (t >= a.nblk) return;                                 // (no block-wide barrier below)

    float *cols = cols_all + (size_t)r * S * 32;
    float *mycol = cols + lane;
    for (int s = 0; s < S; s++) mycol[s * 32] = 0.f;
    for (int s = lane; s < kRowPad; s += 32) {
        float v = 0.f;
        if (s < S) {
            const float w = a.logprob[((size_t)t * a.nbatch + b) * S + s];
            // Z_t = score - offset(alpha_t) - offset(beta_{t+1})
            const float z = score2 - a.coff[(size_t)b * a.nblk + t] -
                            a.coff[((size_t)a.nbatch + b) * a.nblk + t];
            const float sc = s < a.nsharp ? a.sharp * kLog2e : kLog2e;
            v = s < a.ncan ? fmaf(w, sc, -z) : w * sc;
        }
        q[r][s] = v;
    }
    __syncwarp();
    const int off = a.seqoff[b];
    const float *ar = a.fwd_ws + ((size_t)b * a.nblk + t) * Ls;
    const float *br = a.bwd_ws + ((size_t)b * a.nblk + t) * Ls;
    const int32_t *st = a.stayidx + off;
    const int32_t *mv = a.moveidx + (off - b);
    const int32_t *mm = MOD ? a.modmoveidx + (off - b) : nullptr;
    const float *mf = MOD ? a.modmovefact + (off - b) : nullptr;
    const float *qr = q[r];
    for (int p0 = lane; p0 < L; p0 += 4 * 32) {
        float al[4], be[4], be1[4], f[4];
        int is[4], im[4], ix[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {                        // loads of four positions in flight together
            const int p = min(p0 + 32 * u, L - 1);
            const bool has_move = p < L - 1;
            al[u] = __ldcs(ar + p);
            be[u] = __ldcs(br + p);
            be1[u] = has_move ? __ldcs(br + p + 1) : 0.f;
            is[u] = st[p];
            im[u] = has_move ? mv[p] : -1;
            ix[u] = (MOD && has_move) ? mm[p] : 0;
            f[u] = (MOD && has_move) ? mf[p] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (p0 + 32 * u < L) {
                float *cs = mycol + is[u] * 32;
                *cs += ex2f(al[u] + be[u] + qr[is[u]]);
                if (im[u] >= 0) {
                    float x = al[u] + be1[u] + qr[im[u]];
                    if (MOD) x = fmaf(qr[ix[u]], f[u], x);
                    const float e = ex2f(x);
                    float *cm = mycol + im[u] * 32;
                    *cm += e;
                    if (MOD) {                               // c_cat_mod_flipflop.c:465-466
                        float *cx = mycol + ix[u] * 32;
                        *cx = fmaf(e, f[u], *cx);
                    }
                }
            }
        }
    }
    __syncwarp();
    // sum the 32 columns of each transition: lane l rows l and l + 32, rotated start
    float tot[2] = {0.f, 0.f};
#pragma unroll
    for (int r2 = 0; r2 < 2; r2++) {
        const int row = lane + 32 * r2;
        if (row < S) {
#pragma unroll 8
            for (int k = 0; k < 32; k++) tot[r2] += cols[row * 32 + ((k + lane) & 31)];
        }
    }
    // normalise the canonical transitions to sum to one and write the row once
    const float psum = (lane < a.ncan ? tot[0] : 0.f) + (lane + 32 < a.ncan ? tot[1] : 0.f);
    const float scale = a.grad_scale / warp_sum(psum);
    float *g = a.grad_out + ((size_t)t * a.nbatch + b) * S;
    if (lane < S) g[lane] = scale * tot[0];
    if (lane + 32 < S) g[lane + 32] = scale * tot[1];
}

// ---------------------------------------------------------------------------
// flip-flop index build (flipflopfings.py:6-31, ctc.pyx:127-132, :287-292)
__global__ void indices_kernel(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                               int64_t total, int nbase, const int64_t *mod_cats,
                               const int32_t *can_mods_offsets, const float *mod_cat_weights,
                               int32_t *moveidx, int32_t *stayidx, int32_t *seqlen32,
                               int32_t *modmoveidx, float *modmovefact, int32_t *bad_flag) {
    // s_off[b]: first label of chunk b; s_ne[b]: non-empty chunks before b.
    // Move entries are packed as the reference's Python packs them (one per
    // non-final position, ctc.pyx:127-129), i.e. at i - s_ne[b].
    extern __shared__ int64_t s_off[];   // [nbatch + 1] then [nbatch]
    int64_t *s_ne = s_off + nbatch + 1;
    if (threadIdx.x == 0) {
        int64_t acc = 0, ne = 0;
        for (int i = 0; i < nbatch; i++) {
            s_off[i] = acc; s_ne[i] = ne;
            acc += seqlen[i]; ne += seqlen[i] > 0;
        }
        s_off[nbatch] = acc;
    }
    __syncthreads();
    const int nstate = 2 * nbase;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        // chunk of element i: binary search over the prefix sums
        int lo = 0, hi = nbatch;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= i) lo = mid; else hi = mid;
        }
        const int b = lo;
        // labels outside [0, 2 nbase) are clamped so that no kernel gathers outside a score
        // row; the operator layer raises the reference's assertion (ctc.pyx:133-134) from a
        // device flag it reads back with the loss (ctc.check_pending)
        const int64_t q_raw = seqs[i];
        const int q = min(max((int)q_raw, 0), nstate - 1);
        bool bad = q_raw != q;
        stayidx[i] = q + min(q, nbase) * nstate;
        if (i + 1 < s_off[b + 1]) {
            const int qn = min(max((int)seqs[i + 1], 0), nstate - 1);
            const int64_t j = i - s_ne[b];
            moveidx[j] = q + min(qn, nbase) * nstate;
            if (modmoveidx) {
                const int base = qn % nbase;
                const int nmod = can_mods_offsets[base + 1] - can_mods_offsets[base];
                const int64_t mc_raw = mod_cats[i + 1];
                const int mc = min(max((int)mc_raw, 0), nmod - 1);
                bad |= mc_raw != mc;
                const int modseq = can_mods_offsets[base] + mc;
                modmoveidx[j] = nstate * (nbase + 1) + modseq;
                modmovefact[j] = mod_cat_weights[modseq];
            }
        }
        if (bad && bad_flag) atomicOr(bad_flag, 1);
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < nbatch; i += blockDim.x) seqlen32[i] = (int32_t)seqlen[i];
}

// ---------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct CrfWsLayout {
    size_t seqoff, fb, coff, fwd, bwd, total;
    int Ls;
};

static CrfWsLayout crf_layout(int nblk, int nbatch, int max_seqlen, int want_grad) {
    CrfWsLayout w{};
    w.Ls = (int)align_up((size_t)(max_seqlen > 0 ? max_seqlen : 1), 4);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
    w.seqoff = take((size_t)nbatch * sizeof(int));
    w.fb = take((size_t)nbatch * 2 * sizeof(float));
    if (want_grad) {
        // nblk + 1 rows per chunk: the fused kernel also spills the vector both chains meet at
        w.coff = take((size_t)2 * nbatch * (nblk + 1) * sizeof(float));
        w.fwd = take((size_t)nbatch * (nblk + 1) * w.Ls * sizeof(float));
        w.bwd = take((size_t)nbatch * (nblk + 1) * w.Ls * sizeof(float));
    }
    w.total = o;
    return w;
}

template <int P, bool MOD>
static void launch_chain(const CrfArgs &a, int max_seqlen, cudaStream_t s) {
    int threads = (max_seqlen + P - 1) / P;
    threads = (threads + 31) / 32 * 32;
    if (threads < 32) threads = 32;
    threads += 32;       // the transformer warp
    const int grid = a.nchain;
    crf_chain_kernel<P, MOD><<<grid, threads, 0, s>>>(a);
}

int crf_pick_p(int max_seqlen, bool mod);

}  // namespace ty

using namespace ty;

extern "C" size_t ty_crf_flipflop_workspace_bytes(int ntrans, int nblk, int nbatch,
                                                  int max_seqlen, int want_grad) {
    (void)ntrans;
    return crf_layout(nblk, nbatch, max_seqlen, want_grad).total;
}

ty::CrfTuning &ty::crf_tuning() {
    // defaults from the environment, read once per process; ty_crf_tuning() changes them at run time
    static CrfTuning t = [] {
        CrfTuning v;
        const char *e = getenv("TY_CRF_P");
        v.forced_p = e ? atoi(e) : 0;
        e = getenv("TY_CRF_FUSED");
        v.fused = !(e && e[0] == '0');
        return v;
    }();
    return t;
}

static int g_last_path = 0;
extern "C" int ty_crf_last_path(void) { return g_last_path; }

extern "C" void ty_crf_tuning(int forced_p, int fused) {
    if (forced_p >= 0) crf_tuning().forced_p = forced_p;
    if (fused >= 0) crf_tuning().fused = fused != 0;
}

int ty::crf_pick_p(int max_seqlen, bool mod) {
    // positions per thread: 4 keeps a 440-base chunk in four warps (one per
    // SM sub-partition); longer chunks widen the thread block first.  The cat-mod
    // chain (three gathers per position) measures 7 % faster with 8 positions
    // per thread and two DP warps (profiles/r1_microbench_v4.jsonl, tag B).
    const int forced = crf_tuning().forced_p;      // TY_CRF_P / ty_crf_tuning(): tuning override
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) {
        const int cap = forced >= 8 ? 512 : 992;
        if ((max_seqlen + forced - 1) / forced <= cap) return forced;
    }
    if (mod && max_seqlen <= 2048) return 8;
    if (max_seqlen <= 3968) return 4;
    if (max_seqlen <= 8191) return 16;
    return 0;
}

extern "C" int ty_crf_flipflop(const float *logprob, int ntrans, int nblk, int nbatch,
                               const int32_t *moveidx, const int32_t *stayidx,
                               const int32_t *modmoveidx, const float *modmovefact,
                               const int32_t *seqlen, int max_seqlen, float sharp, int nsharp,
                               float score_scale, float *score_out, float grad_scale,
                               float *grad_out, void *workspace, size_t workspace_bytes,
                               void *stream) {
    if (!logprob || !moveidx || !stayidx || !seqlen || !score_out) {
        set_error("ty_crf_flipflop: null pointer");
        return TY_EINVAL;
    }
    if (ntrans <= 0 || ntrans >= kRowPad || nblk <= 0 || nbatch <= 0 || max_seqlen < 0 ||
        nsharp < 0 || nsharp > ntrans) {
        set_error("ty_crf_flipflop: bad shape ntrans=%d nblk=%d nbatch=%d nsharp=%d", ntrans,
                  nblk, nbatch, nsharp);
        return TY_EINVAL;
    }
    if ((modmoveidx == nullptr) != (modmovefact == nullptr)) {
        set_error("ty_crf_flipflop: modmoveidx and modmovefact must both be given");
        return TY_EINVAL;
    }
    const int want_grad = grad_out != nullptr;
    const CrfWsLayout w = crf_layout(nblk, nbatch, max_seqlen, want_grad);
    if (!workspace || workspace_bytes < w.total) {
        set_error("ty_crf_flipflop: workspace %zu < %zu bytes", workspace_bytes, w.total);
        return TY_EWORKSPACE;
    }
    const int P = crf_pick_p(max_seqlen > 0 ? max_seqlen : 1, modmoveidx != nullptr);
    if (P == 0) {
        set_error("ty_crf_flipflop: max_seqlen %d > 8191 unsupported", max_seqlen);
        return TY_EINVAL;
    }
    char *base = static_cast<char *>(workspace);
    CrfArgs a{};
    a.logprob = logprob; a.ntrans = ntrans; a.nblk = nblk; a.nbatch = nbatch;
    a.moveidx = moveidx; a.stayidx = stayidx; a.modmoveidx = modmoveidx;
    a.modmovefact = modmovefact; a.seqlen = seqlen;
    a.sharp = sharp; a.nsharp = nsharp;
    a.ncan = ntrans;
    if (modmoveidx) {   // cat-mod columns follow the 2*nb*(nb+1) flip-flop block (ctc.pyx:262-264)
        int nb = 1;
        while (2 * (nb + 1) * (nb + 2) <= ntrans) nb++;
        a.ncan = 2 * nb * (nb + 1);
    }
    a.score_scale = score_scale; a.score_out = score_out;
    a.grad_scale = grad_scale; a.grad_out = grad_out;
    a.seqoff = reinterpret_cast<int *>(base + w.seqoff);
    a.fb = reinterpret_cast<float *>(base + w.fb);
    a.coff = reinterpret_cast<float *>(base + w.coff);
    a.fwd_ws = reinterpret_cast<float *>(base + w.fwd);
    a.bwd_ws = reinterpret_cast<float *>(base + w.bwd);
    a.Ls = w.Ls;
    a.want_grad = want_grad;
    a.nchain = want_grad ? 2 * nbatch : nbatch;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool mod = modmoveidx != nullptr;
    // gradient wanted and the rows fit in shared memory: chains with the posterior fused in
    if (want_grad && crf_fused_eligible(P, mod, w.Ls, max_seqlen)) {
        g_last_path = 2;
        return launch_crf_fused(a, P, mod, max_seqlen, s);
    }
    g_last_path = want_grad ? 1 : 3;
#define TY_CHAIN(PP)                                           \
    case PP:                                                   \
        if (mod) launch_chain<PP, true>(a, max_seqlen, s);     \
        else launch_chain<PP, false>(a, max_seqlen, s);        \
        break;
    switch (P) {
        TY_CHAIN(1) TY_CHAIN(2) TY_CHAIN(4) TY_CHAIN(8) TY_CHAIN(16)
    }
#undef TY_CHAIN
    int rc = check_launch("crf_chain_kernel");
    if (rc) return rc;
    if (want_grad) {
        // one warp per row; shared memory holds only the [transition][lane] tables of the block's warps
        const size_t smem = (size_t)kPostWarps * ntrans * 32 * sizeof(float);
        const dim3 grid((nblk + kPostWarps - 1) / kPostWarps, nbatch);
        static bool opted = false;
        if (!opted) {
            cudaFuncSetAttribute(crf_post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
            cudaFuncSetAttribute(crf_post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
            opted = true;
        }
        if (mod) crf_post_kernel<true><<<grid, kPostWarps * 32, smem, s>>>(a);
        else crf_post_kernel<false><<<grid, kPostWarps * 32, smem, s>>>(a);
        rc = check_launch("crf_post_kernel");
    }
    return rc;
}

extern "C" int ty_flipflop_indices(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                                   int64_t total, int nbase, const int64_t *mod_cats,
                                   const int32_t *can_mods_offsets, const float *mod_cat_weights,
                                   int32_t *moveidx, int32_t *stayidx, int32_t *seqlen32,
                                   int32_t *modmoveidx, float *modmovefact, void *stream) {
    return ty_flipflop_indices_checked(seqs, seqlen, nbatch, total, nbase, mod_cats, can_mods_offsets,
                                       mod_cat_weights, moveidx, stayidx, seqlen32, modmoveidx,
                                       modmovefact, nullptr, stream);
}

extern "C" int ty_flipflop_indices_checked(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                                           int64_t total, int nbase, const int64_t *mod_cats,
                                           const int32_t *can_mods_offsets,
                                           const float *mod_cat_weights, int32_t *moveidx,
                                           int32_t *stayidx, int32_t *seqlen32, int32_t *modmoveidx,
                                           float *modmovefact, int32_t *bad_flag, void *stream) {
    if (!seqs || !seqlen || !moveidx || !stayidx || !seqlen32 || nbatch <= 0 || nbase <= 0) {
        set_error("ty_flipflop_indices: bad argument");
        return TY_EINVAL;
    }
    if (modmoveidx && (!mod_cats || !can_mods_offsets || !mod_cat_weights || !modmovefact)) {
        set_error("ty_flipflop_indices: incomplete mod arguments");
        return TY_EINVAL;
    }
    const int threads = 256;
    int grid = (int)((total + threads - 1) / threads);
    if (grid < 1) grid = 1;
    if (grid > 1184) grid = 1184;   // 8 x 148 SMs
    const size_t smem = (size_t)(2 * nbatch + 1) * sizeof(int64_t);
    indices_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(
        seqs, seqlen, nbatch, total, nbase, mod_cats, can_mods_offsets, mod_cat_weights,
        moveidx, stayidx, seqlen32, modmoveidx, modmovefact, bad_flag);
    return check_launch("indices_kernel");
}
