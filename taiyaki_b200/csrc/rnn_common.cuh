// rnn_common.cuh -- PTX wrappers, argument block and cell traits shared by the
// recurrent kernels (rnn_fp32.cu: fp32 recurrence behind the gate-major ABI;
// rnn_ws.cu: warp-specialised kernels behind the unit-major ABI).
#pragma once
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

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
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
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
    unsigned zero;         // always 0; opaque to the compiler (see `late`)
    float *dbias;          // bwd: [G*H] += sum over time and chunks of dxproj (may be null)
    const float *bias;     // fwd: [G*H] added to xproj (may be null)
    __nv_bfloat16 *y16;    // fwd: optional bf16 copy of y (operand of the next GEMMs)
    __nv_bfloat16 *dxproj16, *dhn16;   // bwd: write the gradients as bf16 instead of fp32
};

template <int CELL> struct Cell;
template <> struct Cell<kLstm> { static constexpr int G = 4, NS = 5; };   // i f g o | c
template <> struct Cell<kGru> { static constexpr int G = 3, NS = 4; };    // r z n | W_hn h

// ---------------------------------------------------------------------------
// `reserve` layout (what forward keeps for BPTT), both cells:
//     gates [T][N][H][4] fp32   LSTM: i f g o      GRU: r z n (W_hn h)
//     cstate [T][N][H]   fp32   LSTM: c_t          GRU: unused
// One thread owns all gates of a (unit, chunk) cell, so the gates of a cell
// are one 16-byte store / load.
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
        "r"(v), "r"(rbar)
        : "memory");
}
// Loads that must stay where they are written: `volatile` asm is ordered with
// the other volatile asm (st.async, mbarrier), so placing these after the
// exchange keeps them BELOW the consumers of the current step's inputs.  Hoisted
// above them (as the compiler does with plain __ldg) they share a scoreboard
// slot with the older loads and the consumers end up waiting for the new ones.
__device__ __forceinline__ float ld_nc_pinned(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_nc_pinned4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

}  // namespace ty
