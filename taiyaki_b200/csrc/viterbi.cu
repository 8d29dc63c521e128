// viterbi.cu -- best flip-flop path over the 8-state / 40-transition lattice
// (SURVEY 8(f) row 3).  Replaces cupy_extensions/flipflop.py:387-518
// (flipflop_viterbi) and the PyTorch loop taiyaki/decode.py:79-115 with the
// same outputs: fwd [T+1][N][2 nbase] max-scores, traceback [T][N][2 nbase]
// (best predecessor state), path [T+1][N] (states).
// One warp per chunk.  Lane l < 2 nbase owns target state l and keeps its
// max-score in a register; the previous vector is exchanged by shuffles.  Ties
// resolve to the lowest predecessor index (torch.max on CPU).  The traceback is
// followed by lane 0 after the forward sweep; both passes are sequential in T
// and latency bound, like the training chains.
#include "common.cuh"

namespace ty {

constexpr float kVitLarge = 1e30f;     // taiyaki/constants.py:8 LARGE_VAL

template <int NB>
__global__ void __launch_bounds__(128) viterbi_kernel(const float *__restrict__ scores, int T,
                                                      int N, float *__restrict__ fwd,
                                                      int64_t *__restrict__ traceback,
                                                      int64_t *__restrict__ path) {
    constexpr int NS = 2 * NB, S = 2 * NB * (NB + 1);
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const bool owner = lane < NS;
    const bool flip = lane < NB;
    float v = flip ? 0.f : -kVitLarge;             // decode.py:98-99
    if (owner) fwd[(size_t)n * NS + lane] = v;
    const float *row = scores + (size_t)n * S;
    // prefetch one row ahead: flip target l reads scores[l*NS + from], flop target NB+b
    // reads scores[NS*NB + b] (flip b -> flop b) and scores[NS*NB + NB + b] (flop b stays)
    float w[NS];
    auto load = [&](int t, float (&dst)[NS]) {
        const float *r = row + (size_t)t * N * S;
        if (owner) {
            if (flip) {
#pragma unroll
                for (int f = 0; f < NS; f++) dst[f] = r[lane * NS + f];
            } else {
                dst[0] = r[NS * NB + (lane - NB)];
                dst[1] = r[NS * NB + lane];
            }
        }
    };
    if (T > 0) load(0, w);
    for (int t = 0; t < T; t++) {
        float wn[NS];
        if (t + 1 < T) load(t + 1, wn);
        float pv[NS];                  // previous vector, every lane (warp-uniform shuffles)
#pragma unroll
        for (int f = 0; f < NS; f++) pv[f] = __shfl_sync(kFullMask, v, f);
        float best = -3.0e38f;
        int arg = 0;
        if (flip) {
#pragma unroll
            for (int f = 0; f < NS; f++) {
                const float c = pv[f] + w[f];
                if (c > best) { best = c; arg = f; }
            }
        } else if (owner) {
            const int b = lane - NB;
            float c0 = -3.0e38f, c1 = -3.0e38f;
#pragma unroll
            for (int f = 0; f < NB; f++) {             // flip b -> flop b, flop b stays
                if (f == b) { c0 = pv[f] + w[0]; c1 = pv[NB + f] + w[1]; }
            }
            best = c0; arg = b;
            if (c1 > c0) { best = c1; arg = NB + b; }
        }
        if (owner) {
            v = best;
            fwd[((size_t)(t + 1) * N + n) * NS + lane] = best;
            traceback[((size_t)t * N + n) * NS + lane] = arg;
        }
#pragma unroll
        for (int f = 0; f < NS; f++) w[f] = wn[f];
    }
    // ---- final state (first maximum) and traceback ----
    float bv = owner ? v : -3.0e38f;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(kFullMask, bv, o);
        const int oi = __shfl_xor_sync(kFullMask, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        int64_t s = bi;
        path[(size_t)T * N + n] = s;
        for (int t = T - 1; t >= 0; t--) {
            s = traceback[((size_t)t * N + n) * NS + s];
            path[(size_t)t * N + n] = s;
        }
    }
}

}  // namespace ty

using namespace ty;

extern "C" int ty_flipflop_viterbi(const float *scores, int T, int N, int nbase, float *fwd,
                                   int64_t *traceback, int64_t *path, void *stream) {
    if (!scores || !fwd || !traceback || !path || T < 0 || N <= 0) {
        set_error("ty_flipflop_viterbi: bad argument");
        return TY_EINVAL;
    }
    if (nbase != 4) {
        set_error("ty_flipflop_viterbi: only nbase == 4 (40 transitions) is implemented, got %d", nbase);
        return TY_EINVAL;
    }
    viterbi_kernel<4><<<(N + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        scores, T, N, fwd, traceback, path);
    return check_launch("viterbi_kernel");
}
