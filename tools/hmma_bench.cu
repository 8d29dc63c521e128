// hmma_bench.cu -- latency / issue rate of mma.sync.m16n8k16 bf16 on one warp per
// SM sub-partition: NCH independent accumulator chains, round-robin.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NCH>
__global__ void k(int iters, float *out, long long *cyc) {
    float acc[NCH][4];
    uint32_t a[4] = {threadIdx.x, 2, 3, 4};
#pragma unroll
    for (int c = 0; c < NCH; c++) for (int e = 0; e < 4; e++) acc[c][e] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) mma_bf16(acc[c], a, 5u + c, 7u);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) for (int e = 0; e < 4; e++) s += acc[c][e];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NCH> void run(int warps) {
    float *out; long long *cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<NCH><<<1, 32 * warps>>>(iters, out, cyc);
    k<NCH><<<1, 32 * warps>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/CTA=%d chains=%2d : %.2f cycles per HMMA (per warp)\n", warps, NCH, (double)c / iters / NCH);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w); }
    return 0;
}
