// crf_common.cuh -- argument block, constants and step-loop helpers shared by the
// label-constrained CRF kernels (crf_flipflop.cu: chains + posterior kernel;
// crf_fused.cu: chains with the posterior fused in, meeting in the middle).
#pragma once
#include "common.cuh"

namespace ty {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct CrfArgs {
    const float *logprob;
    int ntrans, nblk, nbatch;
    const int32_t *moveidx, *stayidx, *modmoveidx;
    const float *modmovefact;
    const int32_t *seqlen;
    float sharp;
    int nsharp;        // columns < nsharp are multiplied by sharp
    int ncan;          // columns < ncan are stay / move transitions (carry the shift); the rest are cat-mod
    float score_scale;
    float *score_out;
    float grad_scale;
    float *grad_out;
    // workspace
    int *seqoff;       // [nbatch] prefix sums of seqlen
    float *fb;         // [nbatch][2] forward / backward log2-scores
    float *coff;       // [2][nbatch][nblk] accumulated log2 offsets of the stored rows
    float *fwd_ws;     // [nbatch][nblk][Ls] alpha_t         (log2 domain, normalised)
    float *bwd_ws;     // [nbatch][nblk][Ls] beta_{t+1}
    int Ls;
    int want_grad;
    int nchain;        // CTAs of the chain kernel
};

constexpr int kRing = 8;      // cp.async ring slots for raw score rows
constexpr int kDepth = 6;     // rows in flight
constexpr int kRowPad = 64;   // floats per row slot (ntrans <= 63); slot 63 = -1e30 pad
constexpr int kPadSlot = 63;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2(2^x + 2^y)
__device__ __forceinline__ float logaddexp2(float x, float y) {
    return fmaxf(x, y) + lg2f(1.0f + ex2f(-fabsf(x - y)));
}

// ---------------------------------------------------------------------------
// Shared-memory accesses of the chain's step loop use 32-bit shared-window
// addresses computed once and kept opaque: through generic pointers the compiler
// re-derives the window base (S2R SR_CgaCtaId, ~25 cycles of latency) and the
// thread index (S2R SR_TID.X) inside the loop, on the step's dependency chain.
__device__ __forceinline__ float lds_v_f32(unsigned a) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v_f32(unsigned a, float v) {
    asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(a), "f"(v));
}
__device__ __forceinline__ unsigned opaque(unsigned x) {   // keeps a loop-invariant value in a register
    unsigned y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ unsigned pinned_tid() {         // not rematerialised as S2R in the loop
    unsigned t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t;
}

// tuning knobs (A/B timing): positions per thread forced to 1/2/4/8/16 (0 = automatic); fused
// chain + posterior kernel allowed
struct CrfTuning {
    int forced_p = 0;
    bool fused = true;
};
CrfTuning &crf_tuning();

// fused kernel (crf_fused.cu): 0 when the shape is outside its range (the caller then runs
// crf_chain_kernel + crf_post_kernel), else the launch status
bool crf_fused_eligible(int P, bool mod, int Ls, int max_seqlen);
int launch_crf_fused(CrfArgs a, int P, bool mod, int max_seqlen, cudaStream_t s);

}  // namespace ty
