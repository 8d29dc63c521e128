// remap.cu -- best alignment of a label sequence to a matrix of flip-flop transition
// scores: the max-product twin of the training DP, used to (re)map reads to their
// references.  Replaces taiyaki/flipflop_remap.py:6-86
// (map_to_crf_viterbi, a numpy loop over blocks) with the same result: score of the best
// path and path [T+1] of sequence positions, -1 while in the clipping "start" / "end"
// states (localpen).
//
// One CTA per read, any number of reads per launch (ragged T and M through offset
// vectors).  Position scores are fp64 like the reference's (np.full(M, -LARGE_VAL) is
// float64, the fp32 scores are promoted), so maxima, ties and the returned score are
// bit-identical.  The two score vectors, the per-position transition indices (as bytes) and
// a double-buffered score row live in shared memory (M <= ~12.9k positions; longer reads
// run the same loop on a global-memory workspace); one barrier per block of signal; the
// next row is fetched into a register during the step.  Traceback decisions are one byte
// per (block, position), written coalesced, followed at the end by warp 0 through 32-block
// windows.
#include "common.cuh"

namespace ty {

constexpr double kRemapLarge = 1e30;      // taiyaki/constants.py:8 LARGE_VAL

struct RemapArgs {
    const float *scores;        // [sum T][S]
    const int64_t *t_off;       // [nread + 1]
    const int32_t *step_idx;    // [sum (M - 1)]
    const int32_t *stay_idx;    // [sum M]
    const int64_t *m_off;       // [nread + 1]
    const int64_t *tb_off;      // [nread + 1] offsets into tb (sum T * M bytes)
    int S;
    double localpen;
    double *score;              // [nread]
    int32_t *path;              // [sum (T + 1)]
    uint8_t *tb;
    double *dp;                 // [2 * sum M] (global-memory variant only)
    int mp;                     // padded max M (shared-memory variant)
};

template <bool SMEM>
__global__ void __launch_bounds__(1024) remap_kernel(const RemapArgs a) {
    extern __shared__ double remap_smem[];
    const int r = blockIdx.x, tid = threadIdx.x, B = blockDim.x;
    const int T = (int)(a.t_off[r + 1] - a.t_off[r]);
    const int M = (int)(a.m_off[r + 1] - a.m_off[r]);
    const int S = a.S, Sp = (S + 3) & ~3;
    const float *sc = a.scores + a.t_off[r] * S;
    const int32_t *g_stay = a.stay_idx + a.m_off[r];
    const int32_t *g_step = a.step_idx + (a.m_off[r] - r);
    uint8_t *tb = a.tb + a.tb_off[r];
    int32_t *path = a.path + a.t_off[r] + r;
    const double localpen = a.localpen;

    double *buf0, *buf1;
    float *srow;
    uint8_t *s_stay = nullptr, *s_step = nullptr;
    if constexpr (SMEM) {
        buf0 = remap_smem;
        buf1 = buf0 + a.mp;
        srow = reinterpret_cast<float *>(buf1 + a.mp);
        s_stay = reinterpret_cast<uint8_t *>(srow + 2 * Sp);
        s_step = s_stay + a.mp;
        for (int p = tid; p < M; p += B) {
            s_stay[p] = (uint8_t)g_stay[p];
            if (p + 1 < M) s_step[p] = (uint8_t)g_step[p];
        }
    } else {
        buf0 = a.dp + 2 * a.m_off[r];
        buf1 = buf0 + M;
        srow = reinterpret_cast<float *>(remap_smem);
    }
    auto stay_of = [&](int p) -> int { if constexpr (SMEM) return s_stay[p]; else return g_stay[p]; };
    auto step_of = [&](int p) -> int { if constexpr (SMEM) return s_step[p]; else return g_step[p]; };

    for (int p = tid; p < M; p += B) buf0[p] = p == 0 ? 0.0 : -kRemapLarge;     // :31-33
    for (int n = tid; n <= T; n += B) path[n] = -1;                             // :73
    if (tid < S && T > 0) srow[tid] = sc[tid];
    double start_score = 0.0, end_score = -kRemapLarge;                         // :35-37
    int alignment_end = 0;
    __syncthreads();

    for (int n = 0; n < T; n++) {
        const double *prev = (n & 1) ? buf1 : buf0;
        double *cur = (n & 1) ? buf0 : buf1;
        const float *row = srow + (n & 1) * Sp;
        float next = 0.f;
        const bool fetch = tid < S && n + 1 < T;
        if (fetch) next = sc[(size_t)(n + 1) * S + tid];
        uint8_t *tbrow = tb + (size_t)n * M;
        for (int p = tid; p < M; p += B) {
            const double cstay = prev[p] + (double)row[stay_of(p)];             // :50
            double c;
            bool move;
            if (p > 0) {
                const double cstep = prev[p - 1] + (double)row[step_of(p - 1)]; // :53
                c = fmax(cstay, cstep);                                         // :61
                move = cstay < cstep;                                           // :63
            } else {
                const double leave_start = start_score - localpen;             // :56
                start_score = start_score + fmax((double)row[stay_of(0)], -localpen);   // :57
                c = fmax(cstay, start_score);                                   // :62
                move = leave_start > cstay;                                     // :64
            }
            cur[p] = c;
            tbrow[p] = move ? 1 : 0;
        }
        if (tid == 0) {                                                         // :66-71
            const double remain = end_score + fmax((double)row[stay_of(M - 1)], -localpen);
            const double into = prev[M - 1] - localpen;
            if (into > remain) alignment_end = n;
            end_score = fmax(remain, into);
        }
        if (fetch) srow[((n + 1) & 1) * Sp + tid] = next;
        __syncthreads();
    }

    // ---- traceback (:73-85) by warp 0.  The position drops by at most one per block, so the
    // next 32 decisions lie in a 32-row x 33-column window of the table ending at (n, m): the
    // lanes fetch it with independent loads (one memory latency per 32 blocks instead of one
    // per block), lane 0 walks it.
    if (tid < 32) {
        __shared__ uint8_t win[32][36];
        const int lane = tid;
        int n = 0, m = M - 1;
        if (lane == 0) {
            const double *fin = (T & 1) ? buf1 : buf0;
            const double last = fin[M - 1];
            n = last > end_score ? T : alignment_end;                           // :74-79
            a.score[r] = fmax(last, end_score);
        }
        n = __shfl_sync(kFullMask, n, 0);
        while (n >= 0 && m >= 0) {
            const int rn = n - lane;            // row of the decision table this lane fetches
#pragma unroll
            for (int c = 0; c < 33; c++) {
                const int col = m - 32 + c;
                win[lane][c] = (rn > 0 && col >= 0) ? tb[(size_t)(rn - 1) * M + col] : 0;
            }
            __syncwarp();
            if (lane == 0) {
                const int m0 = m;
                for (int j = 0; j < 32 && n >= 0 && m >= 0; j++) {
                    path[n] = m;
                    m -= win[j][m - m0 + 32];
                    n -= 1;
                }
            }
            n = __shfl_sync(kFullMask, n, 0);
            m = __shfl_sync(kFullMask, m, 0);
            __syncwarp();                       // the window is rewritten next round
        }
    }
}

}  // namespace ty

using namespace ty;

extern "C" int ty_flipflop_remap(const float *scores, const int64_t *t_off, const int32_t *step_idx,
                                 const int32_t *stay_idx, const int64_t *m_off,
                                 const int64_t *tb_off, int nread, int S, int max_m,
                                 double localpen, double *score, int32_t *path, uint8_t *tb_ws,
                                 double *dp_ws, void *stream) {
    if (!scores || !t_off || !stay_idx || !m_off || !tb_off || !score || !path || !tb_ws ||
        nread <= 0 || S <= 0 || S > 255 || max_m <= 0 || (max_m > 1 && !step_idx)) {
        set_error("ty_flipflop_remap: bad argument (nread=%d S=%d max_m=%d)", nread, S, max_m);
        return TY_EINVAL;
    }
    RemapArgs a{scores, t_off, step_idx, stay_idx, m_off, tb_off, S, localpen,
                score, path, tb_ws, dp_ws, (max_m + 7) & ~7};
    const int Sp = (S + 3) & ~3;
    int threads = ((max_m + 31) / 32) * 32;
    threads = threads < 64 ? 64 : (threads > 1024 ? 1024 : threads);
    // thread t stages score column t of a row: at least S threads (alphabets of 6+ bases have S > 64)
    if (threads < ((S + 31) / 32) * 32) threads = ((S + 31) / 32) * 32;
    const size_t smem = (size_t)a.mp * 18 + (size_t)2 * Sp * 4;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (smem + 2048 <= 227 * 1024) {          // 2 KB left for the traceback window
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(remap_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return check_launch("remap_kernel (shared-memory attribute)");
        remap_kernel<true><<<nread, threads, smem, s>>>(a);
    } else {
        if (!dp_ws) {
            set_error("ty_flipflop_remap: %d positions need the global-memory workspace dp_ws", max_m);
            return TY_EWORKSPACE;
        }
        remap_kernel<false><<<nread, threads, (size_t)2 * Sp * 4, s>>>(a);
    }
    return check_launch("remap_kernel");
}
