// crf_fused.cu -- label-constrained flip-flop CRF forward / backward with the posterior
// fused into the chains ("meet in the middle"); replaces crf_chain_kernel + crf_post_kernel
// for chunks whose rows fit in shared memory (up to ~1100 positions).  Same mathematics as crf_flipflop.cu
// (c_crf_flipflop.c:43-516, c_cat_mod_flipflop.c:37-582).
//
// One CLUSTER of two CTAs per chunk: rank 0 runs the forward chain, rank 1 the backward
// chain, in lockstep.  With h = ceil(nblk / 2):
//   phase 1  forward consumes rows 0..h-1 and spills alpha_t; backward consumes rows
//            nblk-1..h and spills beta_{t+1} -- HALF of what the two-kernel path spills;
//   middle   both spill the vector they hold (alpha_h, beta_h), one cluster barrier, both
//            compute the total score log2 sum_p 2^(alpha_h[p] + beta_h[p]) -- the normaliser
//            of every posterior row is known from here on;
//   phase 2  forward consumes rows h..nblk-1: the two terms of its recurrence at position p,
//                a = alpha_t[p] + stay_t(p),   b = alpha_t[p-1] + move_t(p-1 -> p),
//            are exactly the posterior exponents of those two lattice edges once
//            beta_{t+1}[p] (spilled by the partner in phase 1) and the normaliser are added;
//            the DP warps drop a, b into a shared-memory ring and carry on, and eight
//            POSTERIOR WARPS of the same CTA (two per SM sub-partition, in the issue slots
//            the latency-bound DP warps leave free) fetch the partner row with cp.async and
//            scatter 2^x into per-LANE columns of a [transition][lane] table in shared
//            memory (position p -> lane p % 32; a lane owns its column, so the adds are
//            plain read-modify-writes: no atomics, no sorting, no divergence, a fixed
//            summation order), sum the 32 columns of each transition, renormalise like the
//            reference's softmax (c_crf_flipflop.c:401) and write the gradient row.  The
//            backward chain does the same for rows h-1..0 against the spilled alpha_t.
// No second kernel, no second pass over the rows: a spilled row is written once and read
// once (L2-resident at training sizes), the score tensor is read once per direction, the
// gradient written once.
#include <type_traits>

#include "crf_common.cuh"

namespace ty {

// posterior warps per CTA: 8 (template parameter PW; partner-row buffers and column tables are per
// warp); a ring slot is always consumed by the same warp (kFRing % PW == 0) -- the parity waits
// below rely on it
constexpr int kFRing = 8;       // ring slots (a, b rows of the DP warps -> posterior warps)

__device__ __forceinline__ uint32_t f_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void f_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "F_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra F_WAIT_DONE;\n"
        "bra F_WAIT_LOOP;\n"
        "F_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void f_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t f_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void f_named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Dynamic shared memory of the fused kernel
constexpr int kBinRows = 64;    // rows of a posterior warp's [transition][lane] table (ntrans < 64; row ntrans = sink)
struct FusedSmem {
    int Ls, kPW;
    __host__ __device__ size_t slot_floats() const { return 2 * (size_t)Ls + 4; }     // a[Ls] b[Ls] meta[4]
    __host__ __device__ size_t off_ring() const { return 0; }
    __host__ __device__ size_t off_part() const { return off_ring() + kFRing * slot_floats() * 4; }
    __host__ __device__ size_t off_tab() const { return off_part() + (size_t)kPW * 2 * Ls * 4; }
    __host__ __device__ size_t off_mf() const { return off_tab() + (size_t)Ls * 4; }
    __host__ __device__ size_t off_cols() const { return off_mf() + (size_t)Ls * 4; }
    __host__ __device__ size_t total() const { return off_cols() + (size_t)kPW * kBinRows * 32 * 4; }
};

template <int P, bool MOD, int DIR, int kPW>
__device__ __forceinline__ void crf_fused_body(const CrfArgs &a, const int b, const int L, const int off,
                                               unsigned char *dyn) {
    __shared__ __align__(16) float raw[kRing][kRowPad];
    __shared__ __align__(16) float tr[2][kRowPad];
    __shared__ __align__(16) float bnd[2][32];
    __shared__ __align__(16) float wmaxs[2][32];
    __shared__ __align__(8) uint64_t bar_full[kFRing], bar_empty[kFRing];
    __shared__ float red_m[32], red_s[32];
    __shared__ float s_end, s_score2;

    const int tid = (int)pinned_tid(), lane = tid & 31, warp = tid >> 5;
    const int nwarps = (int)(blockDim.x >> 5);
    const int ncw = nwarps - 1 - kPW;               // DP warps
    const bool is_tx = warp == ncw;
    const bool is_post = warp > ncw;
    const int nchain = (ncw + 1) * 32;              // threads of the per-step barrier
    const int S = a.ntrans;
    const int nblk = a.nblk;
    const int nrow = nblk + 1;                      // rows per chunk in the spill / offset arrays
    const int Ls = a.Ls;
    const size_t ld = (size_t)a.nbatch * S;
    const int p0 = tid * P;
    const int h = (nblk + 1) / 2;
    const int n1steps = DIR == 0 ? h : nblk - h;    // steps of phase 1
    static_assert(kFRing % kPW == 0, "a ring slot must always be consumed by the same warp");
    const FusedSmem lay{Ls, kPW};
    float *const ring = reinterpret_cast<float *>(dyn + lay.off_ring());
    const unsigned ring_u32 = opaque(f_smem_u32(ring));
    const unsigned slot_bytes = (unsigned)lay.slot_floats() * 4;
    const unsigned full_u32 = opaque(f_smem_u32(&bar_full[0]));
    const unsigned empty_u32 = opaque(f_smem_u32(&bar_empty[0]));

    if (tid == 0) {
        for (int i = 0; i < kFRing; i++) {
            f_mbar_init(&bar_full[i], ncw + 1);
            f_mbar_init(&bar_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float *const my_ws = DIR == 0 ? a.fwd_ws : a.bwd_ws;           // this chain spills here
    const float *const other_ws = DIR == 0 ? a.bwd_ws : a.fwd_ws;  // and reads the partner's rows here
    float *const my_coff = a.coff + ((size_t)DIR * a.nbatch + b) * nrow;
    const float *const other_coff = a.coff + ((size_t)(DIR ^ 1) * a.nbatch + b) * nrow;
    __syncthreads();

    if (is_post) {
        // ================= posterior warps =================
        const int pw = warp - ncw - 1, ptid = tid - (ncw + 1) * 32;
        // per-position transitions: stay | move << 8 | mod << 16 (row offsets of the column table);
        // the move of position p is the edge INTO p (forward) / OUT OF p (backward), i.e. the edge
        // whose term the DP thread of p holds; edges that do not exist point at the sink row
        uint32_t *tab = reinterpret_cast<uint32_t *>(dyn + lay.off_tab());
        float *tmf = reinterpret_cast<float *>(dyn + lay.off_mf());
        for (int p = ptid; p < L; p += kPW * 32) {
            const uint32_t sink = (uint32_t)S;
            uint32_t st = (uint32_t)a.stayidx[off + p], mv = sink, mm = sink;
            float f = 0.f;
            const int q = DIR == 0 ? p - 1 : p;
            if (q >= 0 && q < L - 1) {
                mv = (uint32_t)a.moveidx[off - b + q];
                if (MOD) {
                    mm = (uint32_t)a.modmoveidx[off - b + q];
                    f = a.modmovefact[off - b + q];
                }
            }
            tab[p] = st | (mv << 8) | (mm << 16);
            if (MOD) tmf[p] = f;
        }
        float *cols = reinterpret_cast<float *>(dyn + lay.off_cols()) + (size_t)pw * kBinRows * 32;
        for (int i = lane; i < kBinRows * 32; i += 32) cols[i] = 0.f;
        f_named_bar(2, kPW * 32);
        float *part = reinterpret_cast<float *>(dyn + lay.off_part()) + (size_t)pw * 2 * Ls;
        const int nrows2 = nblk - n1steps;              // rows whose posterior this CTA produces
        auto row_of = [&](int j) { return DIR == 0 ? h + j : h - 1 - j; };
        auto fetch = [&](int j, int buf) {              // partner row of posterior row j -> part[buf]
            if (j < nrows2) {
                const float *src = other_ws + ((size_t)b * nrow + row_of(j)) * Ls;
                float *dst = part + (size_t)buf * Ls;
                for (int i = lane * 4; i < L; i += 128) cp_async16(dst + i, src + i);
            }
            cp_async_commit();
        };
        f_cluster_sync();                               // the partner's phase-1 rows are visible
        fetch(pw, 0);
        float *const mycol = cols + lane;
        int it = 0;
        float ocoff_next = pw < nrows2 ? other_coff[row_of(pw)] : 0.f;   // partner's offset of the row, one row ahead
        for (int j = pw; j < nrows2; j += kPW, it++) {
            const int slot = j & (kFRing - 1), use = j / kFRing, buf = it & 1;
            const int t = row_of(j);
            const float ocoff = ocoff_next;
            if (j + kPW < nrows2) ocoff_next = other_coff[row_of(j + kPW)];
            fetch(j + kPW, buf ^ 1);
            cp_async_wait<1>();
            __syncwarp();
            f_mbar_wait(full_u32 + slot * 8, (uint32_t)use & 1u);
            const float *ab = ring + (size_t)slot * lay.slot_floats();
            const float *pr = part + (size_t)buf * Ls;
            // exponent of an edge = a/b term + partner value - zoff (see the file header)
            const float zoff = s_score2 - ab[2 * Ls] - ocoff;
            for (int p0 = lane; p0 < L; p0 += 4 * 32) {
                uint32_t tb[4];
                float ea[4], eb[4], fm[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int p = min(p0 + 32 * u, L - 1);
                    tb[u] = tab[p];
                    const float base = pr[p] - zoff;
                    ea[u] = ab[p] + base;
                    eb[u] = ab[Ls + p] + base;
                    if (MOD) fm[u] = tmf[p];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool ok = p0 + 32 * u < L;
                    ea[u] = ok ? ex2f(ea[u]) : 0.f;
                    eb[u] = ok ? ex2f(eb[u]) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    // the stay, move and mod transitions of one position are different rows (or the sink,
                    // whose content is never used): their read-modify-writes are issued together
                    float *cs = mycol + (tb[u] & 0xff) * 32, *cm = mycol + ((tb[u] >> 8) & 0xff) * 32;
                    float *cx = mycol + ((tb[u] >> 16) & 0xff) * 32;
                    const float vs = *cs, vm = *cm, vx = MOD ? *cx : 0.f;
                    *cs = vs + ea[u];
                    *cm = vm + eb[u];
                    // c_cat_mod_flipflop.c:465-466: the mod transition gets the move's posterior times its factor
                    if (MOD) *cx = fmaf(eb[u], fm[u], vx);
                }
            }
            __syncwarp();
            if (lane == 0) f_mbar_arrive(empty_u32 + slot * 8);       // slot consumed
            // sum the 32 columns of each transition (lane l: rows l and l + 32; rotated start so that
            // the lanes hit different banks), clearing them for the next row
            float tot[2] = {0.f, 0.f};
#pragma unroll
            for (int r2 = 0; r2 < 2; r2++) {
                const int row = lane + 32 * r2;
                if (row <= S) {
#pragma unroll 8
                    for (int k = 0; k < 32; k++) {
                        float *c = cols + row * 32 + ((k + lane) & 31);
                        tot[r2] += *c;
                        *c = 0.f;
                    }
                }
            }
            // normalise the canonical transitions to sum to one and write the row once
            float psum = (lane < a.ncan ? tot[0] : 0.f) + (lane + 32 < a.ncan ? tot[1] : 0.f);
            const float scale = a.grad_scale / warp_sum(psum);
            float *g = a.grad_out + ((size_t)t * a.nbatch + b) * S;
            if (lane < S) g[lane] = scale * tot[0];
            if (lane + 32 < S) g[lane + 32] = scale * tot[1];
            __syncwarp();
        }
        cp_async_wait<0>();
        f_cluster_sync();                               // end of kernel (scores exchanged)
        return;
    }

    // ================= DP warps + transformer =================
    const unsigned tr_u32 = opaque(f_smem_u32(&tr[0][0]));
    const unsigned raw_u32 = opaque(f_smem_u32(&raw[0][0]) + lane * 4);
    const unsigned bnd_u32 = opaque(f_smem_u32(&bnd[0][0]));
    const unsigned wmaxs_u32 = opaque(f_smem_u32(&wmaxs[0][0]));

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
            const int q = DIR == 0 ? p - 1 : p;
            if (p < L && q >= 0 && q < L - 1) {
                mv[i] = a.moveidx[off - b + q] * 4;
                if (MOD) {
                    mm[i] = a.modmoveidx[off - b + q] * 4;
                    mf[i] = a.modmovefact[off - b + q];
                }
            }
            if (p == (DIR == 0 ? 0 : L - 1)) al[i] = 0.f;
        }
    }

    const bool l0 = lane < S, l1 = lane + 32 < S;
    const float sc0 = (lane < a.nsharp ? a.sharp : 1.0f) * kLog2e;
    const float sc1 = (lane + 32 < a.nsharp ? a.sharp : 1.0f) * kLog2e;
    const bool can0 = lane < a.ncan, can1 = lane + 32 < a.ncan;
    const long long tstep = DIR == 0 ? (long long)ld : -(long long)ld;
    const float *src = a.logprob + (size_t)b * S + (size_t)(DIR == 0 ? 0 : nblk - 1) * ld + lane;
    auto issue_row = [&](int k) {
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
    float coff_run = 0.f;
    float c_cur = 0.f, c_prev = 0.f;
    float part = 0.f;
    float *coff = my_coff + (DIR == 0 ? 0 : nblk - 1);
    if (is_tx) {
#pragma unroll
        for (int k = 0; k < kDepth; k++) issue_row(k);
        wmaxs[0][lane] = lane < ncw ? 0.f : -3.0e38f;
        wmaxs[1][lane] = lane < ncw ? 0.f : -3.0e38f;
        if (lane == 0) { tr[0][kPadSlot] = kNegLarge; tr[1][kPadSlot] = kNegLarge; }
        cp_async_wait<kDepth - 1>();
        transform_row(0, 0, 0.f);
    } else {
        if (DIR == 0) {
            if (lane == 31) bnd[1][warp] = al[P - 1];
        } else {
            if (lane == 0) bnd[1][warp] = al[0];
        }
    }
    float pend_max = 0.f;
    f_named_bar(1, nchain);

    const long long dstep = DIR == 0 ? (long long)Ls : -(long long)Ls;
    float *dst = my_ws + ((size_t)b * nrow + (DIR == 0 ? 0 : nblk - 1)) * Ls + p0;

    auto spill = [&]() {
        if (P % 4 == 0) {
#pragma unroll
            for (int i = 0; i < P; i += 4)
                if (p0 + i < L)
                    *reinterpret_cast<float4 *>(dst + i) = make_float4(al[i], al[i + 1], al[i + 2], al[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < P; i++)
                if (p0 + i < L) dst[i] = al[i];
        }
    };

    // One time step.  PAR = k & 1 (row / boundary buffers addressed with immediates);
    // PHASE 1: spill the vector; PHASE 2: hand the recurrence's two terms to the posterior warps.
    auto step = [&](const int k, auto par_c, auto phase_c) {
        constexpr int PAR = decltype(par_c)::value;
        constexpr int PHASE = decltype(phase_c)::value;
        const int j = k - n1steps;                       // posterior row of this CTA (phase 2)
        const unsigned slot = (unsigned)j & (kFRing - 1);
        if (is_tx) {
            issue_row(k + kDepth);
            const float c_next = warp_max(lds_v_f32(wmaxs_u32 + (PAR ^ 1) * 128 + lane * 4)) - c_cur - c_prev;
            if (PHASE == 1) {
                if (lane == 0) *coff = coff_run;
                coff += DIR == 0 ? 1 : -1;
            }
            if (lane == (k & 31)) part += c_cur;
            coff_run += c_cur;
            if (PHASE == 2) {
                // offset of this row's a / b terms: the vector's accumulated offset plus the row's shift
                if (j >= kFRing) f_mbar_wait(empty_u32 + slot * 8, (uint32_t)(j / kFRing - 1) & 1u);
                if (lane == 0) {
                    sts_v_f32(ring_u32 + slot * slot_bytes + 2 * Ls * 4, coff_run);
                    f_mbar_arrive(full_u32 + slot * 8);
                }
            }
            c_prev = c_cur;
            c_cur = c_next;
            cp_async_wait<kDepth - 1>();
            if (k + 1 < nblk) transform_row(k + 1, PAR ^ 1, c_next);
        } else {
            if (lane == 0) sts_v_f32(wmaxs_u32 + PAR * 128 + warp * 4, pend_max);
            if (PHASE == 1) {
                spill();
                dst += dstep;
            }
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
            float xa[P], xb[P], nw[P];
            float tmax = -3.0e38f;
#pragma unroll
            for (int i = 0; i < P; i++) {
                float other;
                if (DIR == 0) other = (i == 0) ? nb : al[i - 1];
                else other = (i == P - 1) ? nb : al[i + 1];
                xa[i] = al[i] + gs[i];
                xb[i] = other + gm[i];
                nw[i] = logaddexp2(xa[i], xb[i]);
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
            if (PHASE == 2) {
                if (j >= kFRing) f_mbar_wait(empty_u32 + slot * 8, (uint32_t)(j / kFRing - 1) & 1u);
                float *sa = ring + (size_t)slot * lay.slot_floats() + p0;
                if (P % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < P; i += 4) {
                        if (p0 + i < Ls) {
                            *reinterpret_cast<float4 *>(sa + i) = make_float4(xa[i], xa[i + 1], xa[i + 2], xa[i + 3]);
                            *reinterpret_cast<float4 *>(sa + Ls + i) = make_float4(xb[i], xb[i + 1], xb[i + 2], xb[i + 3]);
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < P; i++) {
                        if (p0 + i < Ls) { sa[i] = xa[i]; sa[Ls + i] = xb[i]; }
                    }
                }
                __syncwarp();
                if (lane == 0) f_mbar_arrive(full_u32 + slot * 8);
            }
        }
        f_named_bar(1, nchain);
    };
    using std::integral_constant;
    auto run = [&](int k0, int k1, auto phase_c) {
        int k = k0;
        if ((k & 1) && k < k1) { step(k, integral_constant<int, 1>{}, phase_c); k++; }
        for (; k + 1 < k1; k += 2) {
            step(k, integral_constant<int, 0>{}, phase_c);
            step(k + 1, integral_constant<int, 1>{}, phase_c);
        }
        if (k < k1) step(k, integral_constant<int, 0>{}, phase_c);
    };

    // ---- phase 1 ----
    run(0, n1steps, integral_constant<int, 1>{});
    // ---- middle: spill the vector in hand, meet the partner, total score ----
    if (is_tx) {
        if (lane == 0) *coff = coff_run;
    } else {
        spill();
    }
    f_cluster_sync();
    {
        // forward holds alpha_h (spilled at row h), backward beta_h (spilled at row h-1)
        const int orow = DIR == 0 ? h - 1 : h;
        float m = -3.0e38f, v[P];
        if (!is_tx) {
            const float *o = other_ws + ((size_t)b * nrow + orow) * Ls + p0;
#pragma unroll
            for (int i = 0; i < P; i++) {
                v[i] = p0 + i < L ? al[i] + __ldcg(o + i) : -3.0e38f;
                m = fmaxf(m, v[i]);
            }
            m = warp_max(m);
            if (lane == 0) red_m[warp] = m;
        }
        f_named_bar(1, nchain);
        if (!is_tx) {
            float bm = -3.0e38f;
            for (int w = 0; w < ncw; w++) bm = fmaxf(bm, red_m[w]);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < P; i++) s += p0 + i < L ? ex2f(v[i] - bm) : 0.f;
            s = warp_sum(s);
            if (lane == 0) red_s[warp] = s;
        }
        f_named_bar(1, nchain);
        if (is_tx && lane == 0) {
            float bm = -3.0e38f, s = 0.f;
            for (int w = 0; w < ncw; w++) { bm = fmaxf(bm, red_m[w]); s += red_s[w]; }
            s_score2 = bm + lg2f(s) + coff_run + __ldcg(other_coff + orow);
        }
        // (the transformer's first full[] arrival of phase 2 publishes s_score2 to the posterior warps)
    }
    // ---- phase 2 ----
    run(n1steps, nblk, integral_constant<int, 2>{});

    // c_crf_flipflop.c:131-132 / :234: final position (forward) or first (backward)
    const int pend = DIR == 0 ? L - 1 : 0;
    if (!is_tx && pend >= p0 && pend < p0 + P) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < P; i++)
            if (p0 + i == pend) v = al[i];
        s_end = v;
    }
    f_named_bar(1, nchain);
    if (is_tx) {
        double tot = (double)part;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
        if (lane == 0) a.fb[2 * b + DIR] = (float)(tot + (double)s_end);       // log2 units
    }
    f_cluster_sync();
    if (DIR == 0 && is_tx && lane == 0) {
        // total score = mean of forward and backward (c_crf_flipflop.c:482-491)
        const float score2 = 0.5f * (__ldcg(a.fb + 2 * b) + __ldcg(a.fb + 2 * b + 1));
        a.score_out[b] = a.score_scale * kLn2 * score2;
    }
}

template <int P, bool MOD, int kPW>
__global__ void __launch_bounds__(P <= 4 ? 640 : 544) crf_fused_kernel(const CrfArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ int s_off;
    const int b = blockIdx.x >> 1;
    const int dir = (int)f_cluster_ctarank();

    if (threadIdx.x < 32) {       // prefix sum of seqlen (c_crf_flipflop.c:447-451)
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
    if (L <= 0) {   // c_crf_flipflop.c:269-272, :458-464: no labels -> zero score and gradient
        const int h = (a.nblk + 1) / 2;
        const int t0 = dir == 0 ? h : 0, t1 = dir == 0 ? a.nblk : h;
        for (int i = threadIdx.x; i < (t1 - t0) * a.ntrans; i += blockDim.x) {
            const int t = t0 + i / a.ntrans, s = i % a.ntrans;
            a.grad_out[((size_t)t * a.nbatch + b) * a.ntrans + s] = 0.f;
        }
        if (threadIdx.x == 0 && dir == 0) a.score_out[b] = 0.f;
        return;
    }
    if (dir == 0) crf_fused_body<P, MOD, 0, kPW>(a, b, L, off, dyn);
    else crf_fused_body<P, MOD, 1, kPW>(a, b, L, off, dyn);
}

static int fused_threads(int P, int max_seqlen, int pw) {
    int threads = (max_seqlen + P - 1) / P;
    threads = (threads + 31) / 32 * 32;
    if (threads < 32) threads = 32;
    return threads + 32 + pw * 32;      // + transformer warp + posterior warps
}

// posterior warps the shape runs with (0: outside the fused kernel's range)
static int fused_pick_pw(int P, int Ls, int max_seqlen) {
    if (!crf_tuning().fused || max_seqlen <= 0 || (P != 4 && P != 8)) return 0;
    // eight posterior warps or none: with four (all that fits beyond ~1140 positions) the posterior
    // paces the chain and the kernel pair is faster (tools/microbench.py sweep, nblk 2000: 1.24 ms
    // against 1.03 ms; cat-mod config B 1.55 against 1.44 ms)
    const int pw = 8;
    if (fused_threads(P, max_seqlen, pw) > (P <= 4 ? 640 : 544)) return 0;
    return FusedSmem{Ls, pw}.total() <= 212 * 1024 ? pw : 0;
}

bool crf_fused_eligible(int P, bool mod, int Ls, int max_seqlen) {
    (void)mod;
    return fused_pick_pw(P, Ls, max_seqlen) != 0;
}

template <int P, bool MOD, int PW>
static int launch_fused(const CrfArgs &a, int max_seqlen, cudaStream_t s) {
    const size_t smem = FusedSmem{a.Ls, PW}.total();
    const int threads = fused_threads(P, max_seqlen, PW);
    // opt in to the largest dynamic shared memory once per device (a driver call otherwise paid per launch)
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaSuccess;
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        e = cudaFuncSetAttribute(crf_fused_kernel<P, MOD, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
        if (e != cudaSuccess) {
            set_error("crf_fused_kernel: cudaFuncSetAttribute(%d bytes): %s", 212 * 1024, cudaGetErrorString(e));
            return TY_ECUDA;
        }
        if (dev >= 0 && dev < 64) opted[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * a.nbatch);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, crf_fused_kernel<P, MOD, PW>, a);
    if (e != cudaSuccess) {
        set_error("crf_fused_kernel: launch (%d threads, %zu bytes): %s", threads, smem, cudaGetErrorString(e));
        return TY_ECUDA;
    }
    return check_launch("crf_fused_kernel");
}

int launch_crf_fused(CrfArgs a, int P, bool mod, int max_seqlen, cudaStream_t s) {
    const int pw = fused_pick_pw(P, a.Ls, max_seqlen);
#define TY_FUSED(PP, MM, WW) \
    if (P == PP && mod == MM && pw == WW) return launch_fused<PP, MM, WW>(a, max_seqlen, s);
    TY_FUSED(4, false, 8) TY_FUSED(4, true, 8) TY_FUSED(8, false, 8) TY_FUSED(8, true, 8)
#undef TY_FUSED
    set_error("crf_fused_kernel: P = %d with %d posterior warps not instantiated", P, pw);
    return TY_EINVAL;
}

}  // namespace ty
