// crf_flipflop.cu -- label-constrained flip-flop CRF forward/backward/posterior
// on sm_100a.  Replaces taiyaki/ctc/c_crf_flipflop.c:43-516 and
// c_cat_mod_flipflop.c:37-582 (the MOD template flag adds the per-move
// modified-base term).
//
// Structure (DESIGN.md "CRF kernels"):
//   crf_chain_kernel   one CTA per (chunk, direction).  The sequence positions
//                      of the chunk are striped over the threads, P contiguous
//                      positions per thread held in registers; the time loop is
//                      the only sequential dimension.  Per step: gather the two
//                      transition scores of each position from the 40(45)-float
//                      row staged in shared memory by a cp.async ring, take the
//                      neighbour's alpha by warp shuffle (one shared-memory word
//                      per warp boundary), log-add-exp, one __syncthreads.
//                      The normaliser is a two-step-stale block max so no
//                      reduction sits on the dependency chain; any finite
//                      per-(t,chunk) scalar is a valid normaliser because the
//                      shifts are summed into the score (c_crf_flipflop.c:73-77,
//                      :124).  alpha_t / beta_{t+1} rows are spilled to the HBM
//                      workspace for the posterior kernel.
//   crf_grad_kernel    one CTA per (block, chunk) row: softmax over the 2L-1
//                      joint stay/move scores (c_crf_flipflop.c:372-413) and a
//                      shared-memory histogram into the ntrans bins, written
//                      once with the operator's scale folded in.
#include "common.cuh"

namespace ty {

struct CrfArgs {
    const float *logprob;
    int ntrans, nblk, nbatch;
    const int32_t *moveidx, *stayidx, *modmoveidx;
    const float *modmovefact;
    const int32_t *seqlen;
    float sharp;
    int nsharp;
    float score_scale;
    float *score_out;
    float grad_scale;
    float *grad_out;
    // workspace
    int *seqoff;     // [nbatch] prefix sums of seqlen
    float *fb;       // [nbatch][2] forward / backward log-scores
    float *fwd_ws;   // [nbatch][nblk][Ls] alpha_t
    float *bwd_ws;   // [nbatch][nblk][Ls] beta_{t+1}
    int Ls;
    int want_grad;
};

constexpr int kRing = 8;      // cp.async ring slots for score rows
constexpr int kDepth = 6;     // rows in flight
constexpr int kRowPad = 64;   // floats per ring slot (ntrans <= 64)

template <int P, bool MOD, int DIR>
__device__ __forceinline__ void crf_chain_body(const CrfArgs &a, const int b, const int L,
                                               const int off, float (*rows)[kRowPad],
                                               float (*bnd)[32], float (*wmaxbuf)[32]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarp = blockDim.x >> 5;
    const int nwarp4 = (nwarp + 3) & ~3;
    const int S = a.ntrans;
    const int nblk = a.nblk;
    const size_t ld = (size_t)a.nbatch * S;
    const float *lp = a.logprob + (size_t)b * S;
    const int p0 = tid * P;

    // Per-position transition indices live in registers for the whole chain.
    int st[P], mv[P], mm[P];
    float mf[P];
    float al[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        const int p = p0 + i;
        st[i] = 0; mv[i] = 0; mm[i] = 0; mf[i] = 0.f;
        if (p < L) st[i] = a.stayidx[off + p];
        // DIR 0 (forward): move INTO p from p-1.  DIR 1: move OUT of p to p+1.
        const int q = DIR == 0 ? p - 1 : p;
        if (q >= 0 && q < L - 1) {
            mv[i] = a.moveidx[off - b + q];
            if (MOD) {
                mm[i] = a.modmoveidx[off - b + q];
                mf[i] = a.modmovefact[off - b + q];
            }
        }
        // positions without a move see a neighbour pinned at -1e30, so the
        // log-add-exp below returns the stay term unchanged (no branch needed)
        // c_crf_flipflop.c:113-116 / :220-224 point priors
        al[i] = (p == (DIR == 0 ? 0 : L - 1)) ? 0.f : kNegLarge;
    }

    auto issue_row = [&](int k) {
        if (k < nblk) {
            const int t = DIR == 0 ? k : nblk - 1 - k;
            for (int i = tid; i < S; i += blockDim.x)      // a block may be a single warp
                cp_async4(&rows[k % kRing][i], lp + (size_t)t * ld + i);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < kDepth; k++) issue_row(k);

    if (tid < 32) {
        wmaxbuf[0][tid] = tid < nwarp ? 0.f : -3.0e38f;
        wmaxbuf[1][tid] = tid < nwarp ? 0.f : -3.0e38f;
    }
    // boundary words of the initial vector, read by iteration 0
    if (DIR == 0) {
        if (lane == 31) bnd[1][warp] = al[P - 1];
    } else {
        if (lane == 0) bnd[1][warp] = al[0];
    }
    float pend_max = 0.f;     // max of the vector currently in registers (per warp)
    double csum = 0.0;
    cp_async_wait<kDepth - 1>();
    __syncthreads();

    float *ws = (DIR == 0 ? a.fwd_ws : a.bwd_ws);
    if (a.want_grad) ws += ((size_t)b * nblk) * a.Ls + p0;

    for (int k = 0; k < nblk; k++) {
        const int t = DIR == 0 ? k : nblk - 1 - k;
        issue_row(k + kDepth);

        // normaliser: max of the vector two steps ago (see header comment)
        float c = -3.0e38f;
        {
            const float4 *wm4 = reinterpret_cast<const float4 *>(wmaxbuf[(k + 1) & 1]);
            for (int w = 0; w < nwarp4 / 4; w++) {
                const float4 v = wm4[w];
                c = fmaxf(c, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            }
        }
        if (lane == 0) wmaxbuf[k & 1][warp] = pend_max;

        // spill alpha_t (DIR 0) / beta_{t+1} (DIR 1) for the posterior kernel
        if (a.want_grad) {
            float *dst = ws + (size_t)t * a.Ls;
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
        }

        const float *row = rows[k % kRing];
        // neighbour value across the thread boundary
        float nb;
        if (DIR == 0) {
            nb = __shfl_up_sync(kFullMask, al[P - 1], 1);
            if (lane == 0) nb = warp > 0 ? bnd[(k + 1) & 1][warp - 1] : kNegLarge;
        } else {
            nb = __shfl_down_sync(kFullMask, al[0], 1);
            if (lane == 31) nb = warp + 1 < nwarp ? bnd[(k + 1) & 1][warp + 1] : kNegLarge;
        }

        float nw[P];
        float tmax = -3.0e38f;
#pragma unroll
        for (int i = 0; i < P; i++) {
            const float stay = al[i] + a.sharp * row[st[i]];
            float other;
            if (DIR == 0) other = (i == 0) ? nb : al[i - 1];
            else other = (i == P - 1) ? nb : al[i + 1];
            float wm = a.sharp * row[mv[i]];
            if (MOD) wm = fmaf(row[mm[i]], mf[i], wm);
            float v = logaddexp(stay, other + wm) - c;
            if (p0 + i >= L) v = kNegLarge;
            nw[i] = v;
            tmax = fmaxf(tmax, v);
        }
#pragma unroll
        for (int i = 0; i < P; i++) al[i] = nw[i];

        if (DIR == 0) {
            if (lane == 31) bnd[k & 1][warp] = al[P - 1];
        } else {
            if (lane == 0) bnd[k & 1][warp] = al[0];
        }
        pend_max = warp_max(tmax);
        if (tid == 0) csum += (double)c;

        cp_async_wait<kDepth - 1>();
        __syncthreads();
    }

    // c_crf_flipflop.c:131-132 / :234: final position (forward) or first (backward)
    const int pend = DIR == 0 ? L - 1 : 0;
    __shared__ float s_end;
    if (pend >= p0 && pend < p0 + P) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < P; i++)
            if (p0 + i == pend) v = al[i];
        s_end = v;
    }
    __syncthreads();
    if (tid == 0) {
        const float score = (float)(csum + (double)s_end);
        if (a.want_grad) a.fb[2 * b + DIR] = score;
        else a.score_out[b] = a.score_scale * score;
    }
}

template <int P, bool MOD>
__global__ void __launch_bounds__(P <= 4 ? 1024 : 512) crf_chain_kernel(const CrfArgs a) {
    __shared__ __align__(16) float rows[kRing][kRowPad];
    __shared__ __align__(16) float bnd[2][32];
    __shared__ __align__(16) float wmaxbuf[2][32];
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
    if (dir == 0) crf_chain_body<P, MOD, 0>(a, b, L, off, rows, bnd, wmaxbuf);
    else crf_chain_body<P, MOD, 1>(a, b, L, off, rows, bnd, wmaxbuf);
}

// ---------------------------------------------------------------------------
constexpr int kGradThreads = 128;

__device__ __forceinline__ void online_update(float &m, float &s, float u) {
    if (u > m) {
        s = s * __expf(m - u) + 1.0f;
        m = u;
    } else {
        s += __expf(u - m);
    }
}

template <bool MOD>
__global__ void __launch_bounds__(kGradThreads) crf_grad_kernel(const CrfArgs a) {
    __shared__ float row[kRowPad];
    __shared__ float bins[kRowPad];
    __shared__ float red_m[kGradThreads / 32], red_s[kGradThreads / 32];

    const int tid = threadIdx.x;
    const int t = blockIdx.x / a.nbatch;
    const int b = blockIdx.x - t * a.nbatch;
    const int S = a.ntrans;
    const int L = a.seqlen[b];
    const size_t rowoff = ((size_t)t * a.nbatch + b) * S;

    if (L <= 0) {
        if (tid < S) a.grad_out[rowoff + tid] = 0.f;
        if (t == 0 && tid == 0) a.score_out[b] = 0.f;
        return;
    }
    if (tid < S) {
        const float w = a.logprob[rowoff + tid];
        row[tid] = tid < a.nsharp ? a.sharp * w : w;
        bins[tid] = 0.f;
    }
    if (t == 0 && tid == 0)   // score = 0.5 (F + B), c_crf_flipflop.c:482-491
        a.score_out[b] = a.score_scale * 0.5f * (a.fb[2 * b] + a.fb[2 * b + 1]);
    __syncthreads();

    const int off = a.seqoff[b];
    const float *al = a.fwd_ws + ((size_t)b * a.nblk + t) * a.Ls;
    const float *be = a.bwd_ws + ((size_t)b * a.nblk + t) * a.Ls;
    const int32_t *st = a.stayidx + off;
    const int32_t *mv = a.moveidx + (off - b);
    const int32_t *mm = MOD ? a.modmoveidx + (off - b) : nullptr;
    const float *mf = MOD ? a.modmovefact + (off - b) : nullptr;

    float m = -3.0e38f, s = 0.f;
    for (int p = tid; p < L; p += kGradThreads) {
        const float av = al[p];
        online_update(m, s, av + be[p] + row[st[p]]);
        if (p < L - 1) {
            float wm = row[mv[p]];
            if (MOD) wm = fmaf(row[mm[p]], mf[p], wm);
            online_update(m, s, av + be[p + 1] + wm);
        }
    }
    // block combine of (max, sum)
    const float wm_ = warp_max(m);
    s = warp_sum(s * __expf(m - wm_));
    if ((tid & 31) == 0) { red_m[tid >> 5] = wm_; red_s[tid >> 5] = s; }
    __syncthreads();
    float M = red_m[0];
#pragma unroll
    for (int w = 1; w < kGradThreads / 32; w++) M = fmaxf(M, red_m[w]);
    float Z = 0.f;
#pragma unroll
    for (int w = 0; w < kGradThreads / 32; w++) Z += red_s[w] * __expf(red_m[w] - M);
    const float invZ = 1.0f / Z;

    for (int p = tid; p < L; p += kGradThreads) {
        const float av = al[p];
        const int si = st[p];
        atomicAdd(&bins[si], __expf(av + be[p] + row[si] - M) * invZ);
        if (p < L - 1) {
            const int mi = mv[p];
            float wm = row[mi];
            int mmi = 0; float f = 0.f;
            if (MOD) { mmi = mm[p]; f = mf[p]; wm = fmaf(row[mmi], f, wm); }
            const float pm = __expf(av + be[p + 1] + wm - M) * invZ;
            atomicAdd(&bins[mi], pm);
            if (MOD) atomicAdd(&bins[mmi], pm * f);   // c_cat_mod_flipflop.c:465-466
        }
    }
    __syncthreads();
    if (tid < S) a.grad_out[rowoff + tid] = a.grad_scale * bins[tid];
}

// ---------------------------------------------------------------------------
// flip-flop index build (flipflopfings.py:6-31, ctc.pyx:127-132, :287-292)
__global__ void indices_kernel(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                               int64_t total, int nbase, const int64_t *mod_cats,
                               const int32_t *can_mods_offsets, const float *mod_cat_weights,
                               int32_t *moveidx, int32_t *stayidx, int32_t *seqlen32,
                               int32_t *modmoveidx, float *modmovefact) {
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
        const int q = (int)seqs[i];
        stayidx[i] = q + min(q, nbase) * nstate;
        if (i + 1 < s_off[b + 1]) {
            const int qn = (int)seqs[i + 1];
            const int64_t j = i - s_ne[b];
            moveidx[j] = q + min(qn, nbase) * nstate;
            if (modmoveidx) {
                const int modseq = can_mods_offsets[qn % nbase] + (int)mod_cats[i + 1];
                modmoveidx[j] = nstate * (nbase + 1) + modseq;
                modmovefact[j] = mod_cat_weights[modseq];
            }
        }
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < nbatch; i += blockDim.x) seqlen32[i] = (int32_t)seqlen[i];
}

// ---------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct CrfWsLayout {
    size_t seqoff, fb, fwd, bwd, total;
    int Ls;
};

static CrfWsLayout crf_layout(int nblk, int nbatch, int max_seqlen, int want_grad) {
    CrfWsLayout w{};
    w.Ls = (int)align_up((size_t)(max_seqlen > 0 ? max_seqlen : 1), 4);
    size_t o = 0;
    w.seqoff = o; o += align_up((size_t)nbatch * sizeof(int), 256);
    w.fb = o; o += align_up((size_t)nbatch * 2 * sizeof(float), 256);
    const size_t mat = want_grad ? align_up((size_t)nbatch * nblk * w.Ls * sizeof(float), 256) : 0;
    w.fwd = o; o += mat;
    w.bwd = o; o += mat;
    w.total = o;
    return w;
}

template <int P, bool MOD>
static void launch_chain(const CrfArgs &a, int max_seqlen, cudaStream_t s) {
    int threads = (max_seqlen + P - 1) / P;
    threads = (threads + 31) / 32 * 32;
    if (threads < 32) threads = 32;
    const int grid = a.want_grad ? 2 * a.nbatch : a.nbatch;
    crf_chain_kernel<P, MOD><<<grid, threads, 0, s>>>(a);
}

int crf_pick_p(int max_seqlen);

}  // namespace ty

using namespace ty;

extern "C" size_t ty_crf_flipflop_workspace_bytes(int ntrans, int nblk, int nbatch,
                                                  int max_seqlen, int want_grad) {
    (void)ntrans;
    return crf_layout(nblk, nbatch, max_seqlen, want_grad).total;
}

int ty::crf_pick_p(int max_seqlen) {
    // positions per thread: 4 keeps a 440-base chunk in four warps (one per
    // SM sub-partition); longer chunks widen the thread block first.
    const char *e = getenv("TY_CRF_P");   // tuning override, read per call
    const int forced = e ? atoi(e) : 0;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) {
        const int cap = forced >= 8 ? 512 : 1024;
        if ((max_seqlen + forced - 1) / forced <= cap) return forced;
    }
    if (max_seqlen <= 4096) return 4;
    if (max_seqlen <= 8192) return 16;
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
    if (ntrans <= 0 || ntrans > kRowPad || nblk <= 0 || nbatch <= 0 || max_seqlen < 0) {
        set_error("ty_crf_flipflop: bad shape ntrans=%d nblk=%d nbatch=%d", ntrans, nblk, nbatch);
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
    const int P = crf_pick_p(max_seqlen > 0 ? max_seqlen : 1);
    if (P == 0) {
        set_error("ty_crf_flipflop: max_seqlen %d > 8192 unsupported", max_seqlen);
        return TY_EINVAL;
    }
    char *base = static_cast<char *>(workspace);
    CrfArgs a{};
    a.logprob = logprob; a.ntrans = ntrans; a.nblk = nblk; a.nbatch = nbatch;
    a.moveidx = moveidx; a.stayidx = stayidx; a.modmoveidx = modmoveidx;
    a.modmovefact = modmovefact; a.seqlen = seqlen;
    a.sharp = sharp; a.nsharp = nsharp;
    a.score_scale = score_scale; a.score_out = score_out;
    a.grad_scale = grad_scale; a.grad_out = grad_out;
    a.seqoff = reinterpret_cast<int *>(base + w.seqoff);
    a.fb = reinterpret_cast<float *>(base + w.fb);
    a.fwd_ws = reinterpret_cast<float *>(base + w.fwd);
    a.bwd_ws = reinterpret_cast<float *>(base + w.bwd);
    a.Ls = w.Ls;
    a.want_grad = want_grad;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool mod = modmoveidx != nullptr;
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
        const int grid = nblk * nbatch;
        if (mod) crf_grad_kernel<true><<<grid, kGradThreads, 0, s>>>(a);
        else crf_grad_kernel<false><<<grid, kGradThreads, 0, s>>>(a);
        rc = check_launch("crf_grad_kernel");
    }
    return rc;
}

extern "C" int ty_flipflop_indices(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                                   int64_t total, int nbase, const int64_t *mod_cats,
                                   const int32_t *can_mods_offsets, const float *mod_cat_weights,
                                   int32_t *moveidx, int32_t *stayidx, int32_t *seqlen32,
                                   int32_t *modmoveidx, float *modmovefact, void *stream) {
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
        moveidx, stayidx, seqlen32, modmoveidx, modmovefact);
    return check_launch("indices_kernel");
}
