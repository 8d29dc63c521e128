// logz.cu -- partition function of the flip-flop CRF over the 8-state /
// 40-transition lattice and its gradient (posterior transition probabilities).
// Replaces cupy_extensions/flipflop.py:10-368 (flipflop_fwd, flipflop_bwd,
// flipflop_make_trans, LogZ) and the TorchScript fallback layers.py:1253-1299.
//
//   logz_chain_kernel  one WARP per (chunk, direction); lane l owns flip-target
//                      transition l (to = l>>3, from = l&7), lanes 0..7 also own
//                      flop-target transition 32+l.  The state vector is kept
//                      replicated in registers; a step is exp -> 3 (2) xor
//                      shuffles -> 1 gather shuffle -> log.  The shift inside
//                      the log-sum-exp is the row maximum, computed when the
//                      row is prefetched, and the normaliser is the maximum of
//                      the previous state vector, so neither reduction is on
//                      the dependency chain.  Rows are prefetched eight steps
//                      ahead in registers.
//   logz_post_kernel   one warp per (block, chunk) row: softmax over the 40
//                      transitions of fwd[t,from] + w + bwd[t+1,to]
//                      (cupy_extensions/flipflop.py:280-294, :352-354).
#include <math.h>

#include "common.cuh"

namespace ty {

struct LogzArgs {
    const float *scores;
    int ld, nblk, nbatch;
    float logz_scale;
    float *logz_out;
    float grad_scale;
    float *grad_out;
    int ld_grad;
    int accumulate;
    float *fwdz;   // [nblk][nbatch][8]  phi_t (normalised)
    float *bwdz;   // [nblk][nbatch][8]  psi_{t+1} (normalised)
    int want_grad;
};

constexpr int kU = 8;                      // prefetch distance (steps)
constexpr float kFlopInit = -50000.0f;     // layers.py:1289 (LARGE_LOG_VAL)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// The chains run in the log2 domain (scores scaled by log2 e when loaded) so a
// step is bare ex2 / lg2 SFU operations; the stored lattice vectors are log2.
__device__ __forceinline__ float ex2f_(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f_(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float max8(float v) {   // max over aligned groups of 8 lanes
    v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 1));
    v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 2));
    v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 4));
    return v;
}

__global__ void __launch_bounds__(256) logz_chain_kernel(const LogzArgs a) {
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nchain = a.want_grad ? 2 * a.nbatch : a.nbatch;
    if (chain >= nchain) return;
    const int b = a.want_grad ? chain >> 1 : chain;
    const int dir = a.want_grad ? chain & 1 : 0;
    const int nblk = a.nblk;
    const size_t ldt = (size_t)a.nbatch * a.ld;
    const float *w = a.scores + (size_t)b * a.ld;
    const int from = lane & 7, to = lane >> 3;

    float c1[kU], c2[kU], n1[kU], n2[kU];
    auto load = [&](float *r1, float *r2, int k0) {
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const int k = k0 + u;
            if (k < nblk) {
                const int t = dir == 0 ? k : nblk - 1 - k;
                // raw scores: they are scaled by log2(e) where they are consumed, so the
                // loaded registers are not touched until then (keeps the prefetch distance)
                r1[u] = __ldg(w + (size_t)t * ldt + lane);
                r2[u] = __ldg(w + (size_t)t * ldt + 32 + from);
            } else {
                r1[u] = 0.f; r2[u] = 0.f;
            }
        }
    };
    load(c1, c2, 0);

    double acc = 0.0;
    if (dir == 0) {
        // phi: lane holds phi[from]
        float phi = from < 4 ? 0.f : kFlopInit;
        float c = 0.f;    // max of the current (normalised) state vector
        for (int k0 = 0; k0 < nblk; k0 += kU) {
            load(n1, n2, k0 + kU);
            float m[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) m[u] = kLog2e * warp_max(fmaxf(c1[u], c2[u]));
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int k = k0 + u;
                if (k < nblk) {
                    if (a.want_grad && lane < 8)
                        a.fwdz[((size_t)k * a.nbatch + b) * 8 + lane] = phi;
                    const float base = phi - m[u];
                    float e1 = ex2f_(fmaf(kLog2e, c1[u], base));
                    float e2 = ex2f_(fmaf(kLog2e, c2[u], base));
                    e1 += __shfl_xor_sync(kFullMask, e1, 1);
                    e2 += __shfl_xor_sync(kFullMask, e2, 4);
                    e1 += __shfl_xor_sync(kFullMask, e1, 2);
                    e1 += __shfl_xor_sync(kFullMask, e1, 4);
                    // lane needs the new value of state `from`
                    const float g1 = __shfl_sync(kFullMask, e1, (lane & 3) << 3);
                    const float g2 = __shfl_sync(kFullMask, e2, lane & 3);
                    const float raw = lg2f_(from < 4 ? g1 : g2);
                    phi = raw - c;
                    if (lane == 0) acc += (double)m[u] + (double)c;
                    c = max8(raw) - c;
                }
            }
#pragma unroll
            for (int u = 0; u < kU; u++) { c1[u] = n1[u]; c2[u] = n2[u]; }
        }
        // free end: logZ = offsets + logsumexp over the 8 states
        float e = ex2f_(phi - c);
        e += __shfl_xor_sync(kFullMask, e, 1);
        e += __shfl_xor_sync(kFullMask, e, 2);
        e += __shfl_xor_sync(kFullMask, e, 4);
        if (lane == 0)
            a.logz_out[b] = a.logz_scale * kLn2 * (float)(acc + (double)c + (double)lg2f_(e));
    } else {
        // psi: lane holds pa = psi[to] and pb = psi[4 + (from & 3)]
        const float init = -3.0f;         // -log2(8): cupy_extensions/flipflop.py:166
        float pa = init, pb = init;
        float c = init;
        for (int k0 = 0; k0 < nblk; k0 += kU) {
            load(n1, n2, k0 + kU);
            float m[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) m[u] = kLog2e * warp_max(fmaxf(c1[u], c2[u]));
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int k = k0 + u;
                if (k < nblk) {
                    const int t = nblk - 1 - k;
                    if (a.want_grad) {
                        // lanes 0,8,16,24 hold psi[0..3] in pa; lanes 0..3 hold psi[4..7] in pb
                        if ((lane & 7) == 0)
                            a.bwdz[((size_t)t * a.nbatch + b) * 8 + to] = pa;
                        if (lane < 4)
                            a.bwdz[((size_t)t * a.nbatch + b) * 8 + 4 + lane] = pb;
                    }
                    float e1 = ex2f_(fmaf(kLog2e, c1[u], pa - m[u]));
                    const float e2 = ex2f_(fmaf(kLog2e, c2[u], pb - m[u]));
                    e1 += __shfl_xor_sync(kFullMask, e1, 8);
                    e1 += __shfl_xor_sync(kFullMask, e1, 16);
                    const float raw = lg2f_(e1 + e2);       // new psi[from]
                    const float ga = __shfl_sync(kFullMask, raw, to);
                    const float gb = __shfl_sync(kFullMask, raw, 4 + (lane & 3));
                    pa = ga - c;
                    pb = gb - c;
                    c = max8(raw) - c;
                }
            }
#pragma unroll
            for (int u = 0; u < kU; u++) { c1[u] = n1[u]; c2[u] = n2[u]; }
        }
    }
}

__global__ void __launch_bounds__(256) logz_post_kernel(const LogzArgs a) {
    const int lane = threadIdx.x & 31;
    const size_t rowi = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nrow = (size_t)a.nblk * a.nbatch;
    if (rowi >= nrow) return;
    const float *w = a.scores + rowi * a.ld;
    const float *f = a.fwdz + rowi * 8;
    const float *p = a.bwdz + rowi * 8;
    const int from = lane & 7, to = lane >> 3;
    const float phi = f[from];
    const float x1 = fmaf(kLog2e, w[lane], phi + p[to]);
    const float x2 = lane < 8 ? fmaf(kLog2e, w[32 + lane], phi + p[4 + (lane & 3)]) : -3.0e38f;
    const float M = warp_max(fmaxf(x1, x2));
    const float e1 = ex2f_(x1 - M);
    const float e2 = lane < 8 ? ex2f_(x2 - M) : 0.f;
    const float Z = warp_sum(e1 + e2);
    const float sc = a.grad_scale / Z;
    float *g = a.grad_out + rowi * a.ld_grad;
    if (a.accumulate) {
        g[lane] += sc * e1;
        if (lane < 8) g[32 + lane] += sc * e2;
    } else {
        g[lane] = sc * e1;
        if (lane < 8) g[32 + lane] = sc * e2;
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace ty

using namespace ty;

extern "C" size_t ty_flipflop_logz_workspace_bytes(int nbase, int nblk, int nbatch) {
    (void)nbase;
    return 2 * align_up((size_t)nblk * nbatch * 8 * sizeof(float), 256);
}

// phases: bit 0 = lattice chains (logZ and, with grad_out, the stored vectors),
//         bit 1 = posterior (gradient) from the stored vectors,
//         bit 2 = the chains run beside other latency-bound kernels: give them SMs of their own
extern "C" int ty_flipflop_logz_phase(const float *scores, int ld, int nblk, int nbatch,
                                      int nbase, float logz_scale, float *logz_out,
                                      float grad_scale, float *grad_out, int ld_grad,
                                      int accumulate, void *workspace, size_t workspace_bytes,
                                      int phases, void *stream) {
    if (!scores || !logz_out || nblk <= 0 || nbatch <= 0) {
        set_error("ty_flipflop_logz: bad argument");
        return TY_EINVAL;
    }
    if (nbase != 4) {
        set_error("ty_flipflop_logz: only nbase == 4 (40 transitions) is implemented, got %d", nbase);
        return TY_EINVAL;
    }
    if (ld < 40 || (grad_out && ld_grad < 40)) {
        set_error("ty_flipflop_logz: row stride < 40");
        return TY_EINVAL;
    }
    const int want_grad = grad_out != nullptr;
    const size_t half = align_up((size_t)nblk * nbatch * 8 * sizeof(float), 256);
    if (want_grad && (!workspace || workspace_bytes < 2 * half)) {
        set_error("ty_flipflop_logz: workspace %zu < %zu bytes", workspace_bytes, 2 * half);
        return TY_EWORKSPACE;
    }
    LogzArgs a{};
    a.scores = scores; a.ld = ld; a.nblk = nblk; a.nbatch = nbatch;
    a.logz_scale = logz_scale; a.logz_out = logz_out;
    a.grad_scale = grad_scale; a.grad_out = grad_out; a.ld_grad = ld_grad;
    a.accumulate = accumulate;
    a.fwdz = static_cast<float *>(workspace);
    a.bwdz = reinterpret_cast<float *>(static_cast<char *>(workspace) + half);
    a.want_grad = want_grad;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int nchain = want_grad ? 2 * nbatch : nbatch;
    int rc = TY_OK;
    if (phases & 1) {
        if (phases & 4) {
            // Running beside the label-constrained chains (ty_flipflop_train_loss): 8 chains
            // per CTA and a shared-memory reservation no other CTA fits next to, so these
            // warps get SMs of their own instead of sharing a sub-partition with a DP warp
            // of crf_chain_kernel (measured on the cat-mod shape: 2.04 -> see profiles).
            static bool attr = false;
            const int reserve = 200 * 1024;
            if (!attr) {
                cudaFuncSetAttribute(logz_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, reserve);
                attr = true;
            }
            logz_chain_kernel<<<(nchain + 7) / 8, 256, reserve, s>>>(a);
        } else {
            // one chain (warp) per CTA spreads the latency-bound chains over the SMs
            const int warps_per_cta = nchain <= 148 * 4 ? 1 : 4;
            const int grid = (nchain + warps_per_cta - 1) / warps_per_cta;
            logz_chain_kernel<<<grid, 32 * warps_per_cta, 0, s>>>(a);
        }
        rc = check_launch("logz_chain_kernel");
    }
    if (rc || !want_grad || !(phases & 2)) return rc;
    const size_t nrow = (size_t)nblk * nbatch;
    logz_post_kernel<<<(unsigned)((nrow + 7) / 8), 256, 0, s>>>(a);
    return check_launch("logz_post_kernel");
}

extern "C" int ty_flipflop_logz(const float *scores, int ld, int nblk, int nbatch, int nbase,
                                float logz_scale, float *logz_out, float grad_scale,
                                float *grad_out, int ld_grad, int accumulate, void *workspace,
                                size_t workspace_bytes, void *stream) {
    return ty_flipflop_logz_phase(scores, ld, nblk, nbatch, nbase, logz_scale, logz_out,
                                  grad_scale, grad_out, ld_grad, accumulate, workspace,
                                  workspace_bytes, 3, stream);
}
