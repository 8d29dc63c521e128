// common.cuh -- shared device helpers for libtaiyaki_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/taiyaki_b200.h"

namespace ty {

constexpr float kNegLarge = -1e30f;       // the reference's LARGE_VAL (c_crf_flipflop.c:11)
constexpr unsigned kFullMask = 0xffffffffu;

void set_error(const char *fmt, ...);
int check_launch(const char *what);

// max(x,y) + log(1 + exp(-|x-y|))   (vect_mathfun.h:79-102) on the SFU path:
// ex2.approx / lg2.approx carry ~2^-22 absolute error on [0, log 2], far
// inside the 1e-4 parity budget.
__device__ __forceinline__ float logaddexp(float x, float y) {
    const float mx = fmaxf(x, y);
    const float d = -fabsf(x - y);
    return mx + __logf(1.0f + __expf(d));
}

// Order-preserving float <-> int map so warp max can use one REDUX.
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_max(float v) {
    return ordered_to_float(__reduce_max_sync(kFullMask, float_to_ordered(v)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

}  // namespace ty
