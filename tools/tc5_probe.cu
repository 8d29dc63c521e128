// tc5_probe.cu -- what the recurrence kernel needs to know about tcgen05 on this part,
// measured instead of assumed (tools, not product):
//   P1  packing of a bf16 A operand held in tensor memory (which half of a column is even k)
//   P2  roles of the leading / stride byte offsets for an MN-major, un-swizzled B tile
//       ([unit][8 chunks] rows of 16 bytes -- the layout h_t is exchanged in), and whether
//       a zero stride (second 8-column group = the first) is honoured
//   P3  thread <-> element map of tcgen05.ld.16x256b and a lane base of 32w + 16
//   P4  latency of one recurrent product issued as tcgen05.mma (M = 128, N = 16, K = H,
//       A in tensor memory): first issue -> commit observed -> accumulators in registers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build_variants/tc5_probe tools/tc5_probe.cu
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../taiyaki_b200/csrc/tc5.cuh"

using namespace tc5;

constexpr int H = 256, NB = 8;

struct Opt {
    int swap_half;   // 1: odd k in the low half of a tensor-memory column
    int swap_lbo;    // 1: LBO = N-direction stride, SBO = K-direction stride
    int sbo_zero;    // 1: N-direction stride 0 (duplicate the 8 real columns)
    int reps;
    int nacc;        // independent accumulators the K = H product is spread over (1, 4, 16)
    int nmma;        // MMAs issued (16 = the whole product)
};

__global__ void __launch_bounds__(160, 1) probe(const float *W, const float *h, float *D32, float *D16,
                                                long long *clk, Opt o) {
    __shared__ __align__(1024) unsigned char smem[2 * H * 16 + 64];
    unsigned char *hs = smem, *pad = smem + H * 16;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * H * 16);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(smem + 2 * H * 16 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2 * H * 16 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    __syncthreads();
    // B: h[unit][chunk] bf16
    for (int i = tid; i < H * NB; i += blockDim.x) {
        const int u = i / NB, c = i % NB;
        reinterpret_cast<__nv_bfloat16 *>(hs)[u * NB + c] = __float2bfloat16(h[u * NB + c]);
    }
    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<256>(smem_u32(tptr));
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tptr;
    // A: row = tid (0..127) -> lane tid; column c holds k = 2c, 2c+1
    if (warp < 4) {
        const float *src = W + (size_t)tid * H;
        for (int kc = 0; kc < H / 16; kc++) {
            uint32_t v[8];
            for (int j = 0; j < 8; j++) {
                const float lo = src[kc * 16 + 2 * j], hi = src[kc * 16 + 2 * j + 1];
                __nv_bfloat162 p = o.swap_half ? __floats2bfloat162_rn(hi, lo) : __floats2bfloat162_rn(lo, hi);
                v[j] = *reinterpret_cast<uint32_t *>(&p);
            }
            tmem_st_32x8(tmem + ((uint32_t)(warp * 32) << 16) + kc * 8, v);
        }
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t dcol = 128;
    const uint32_t kdir = 128, ndir = o.sbo_zero ? 0u : (uint32_t)(pad - hs);
    const uint32_t lbo = o.swap_lbo ? ndir : kdir, sbo = o.swap_lbo ? kdir : ndir;
    const uint64_t bdesc = (uint64_t(1) << 46) | (uint64_t(sbo >> 4) << 32) | (uint64_t(lbo >> 4) << 16) |
                           uint64_t((smem_u32(hs) & 0x3FFFF) >> 4);
    const uint32_t idesc = idesc_bf16(128, 16, false, true);
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int rep = 0; rep < o.reps; rep++) {
        if (tid == 0) {
            t0 = clock64();
            for (int k = 0; k < o.nmma; k++)
                mma_ts(tmem + dcol + 16 * (k % o.nacc), tmem + k * 8, bdesc + (uint64_t)((256 * k) >> 4), idesc,
                       k >= o.nacc);
            t3 = clock64();
            mma_commit(smem_u32(bar));
        }
        mbar_wait(smem_u32(bar), rep & 1);
        fence_after_sync();
        if (tid == 0) t1 = clock64();
        if (warp < 4) {
            uint32_t a[4], b[4];
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                         : "r"(tmem + ((uint32_t)(warp * 32) << 16) + dcol));
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                         : "r"(tmem + ((uint32_t)(warp * 32 + 16) << 16) + dcol));
            tmem_ld_wait();
            if (tid == 0) t2 = clock64();
            // assumed map: a0,a1 = (lane t/4, cols 2(t%4), +1); a2,a3 = lane t/4 + 8; b = +16
            const int r = lane >> 2, q = lane & 3;
            const int rows[4] = {warp * 32 + r, warp * 32 + r + 8, warp * 32 + 16 + r, warp * 32 + 24 + r};
            D16[rows[0] * 16 + 2 * q] = __uint_as_float(a[0]);
            D16[rows[0] * 16 + 2 * q + 1] = __uint_as_float(a[1]);
            D16[rows[1] * 16 + 2 * q] = __uint_as_float(a[2]);
            D16[rows[1] * 16 + 2 * q + 1] = __uint_as_float(a[3]);
            D16[rows[2] * 16 + 2 * q] = __uint_as_float(b[0]);
            D16[rows[2] * 16 + 2 * q + 1] = __uint_as_float(b[1]);
            D16[rows[3] * 16 + 2 * q] = __uint_as_float(b[2]);
            D16[rows[3] * 16 + 2 * q + 1] = __uint_as_float(b[3]);
            uint32_t v[16];
            tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + dcol, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; j++) D32[tid * 16 + j] = __uint_as_float(v[j]);
            fence_before_sync();
        }
        __syncthreads();
        fence_after_sync();
        if (tid == 0 && rep == o.reps - 1) {
            clk[0] = t1 - t0;
            clk[1] = t2 - t0;
            clk[2] = t3 - t0;
        }
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc<256>(tmem);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    std::vector<float> W(128 * H), h(H * NB), ref(128 * NB);
    srand(1);
    for (auto &x : W) x = bf((rand() / (float)RAND_MAX - 0.5f));
    for (auto &x : h) x = bf((rand() / (float)RAND_MAX - 0.5f));
    for (int r = 0; r < 128; r++)
        for (int c = 0; c < NB; c++) {
            double s = 0;
            for (int k = 0; k < H; k++) s += (double)W[r * H + k] * h[k * NB + c];
            ref[r * NB + c] = (float)s;
        }
    float *dW, *dh, *d32, *d16;
    long long *dclk;
    cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dh, h.size() * 4);
    cudaMalloc(&d32, 128 * 16 * 4); cudaMalloc(&d16, 128 * 16 * 4); cudaMalloc(&dclk, 32);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int combo = 0; combo < 8; combo++) {
        Opt o{combo & 1, (combo >> 1) & 1, (combo >> 2) & 1, 20, 1, 16};
        cudaMemset(d32, 0, 128 * 16 * 4); cudaMemset(d16, 0, 128 * 16 * 4);
        probe<<<1, 160>>>(dW, dh, d32, d16, dclk, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("combo %d: CUDA error %s\n", combo, cudaGetErrorString(e)); return 1; }
        std::vector<float> o32(128 * 16), o16(128 * 16);
        long long clk[3];
        cudaMemcpy(o32.data(), d32, o32.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(o16.data(), d16, o16.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(clk, dclk, 24, cudaMemcpyDeviceToHost);
        double e32 = 0, edup = 0, e16 = 0;
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < NB; c++) {
                e32 = fmax(e32, fabs(o32[r * 16 + c] - ref[r * NB + c]));
                edup = fmax(edup, fabs(o32[r * 16 + 8 + c] - (o.sbo_zero ? ref[r * NB + c] : 0.0f)));
                e16 = fmax(e16, fabs(o16[r * 16 + c] - o32[r * 16 + c]));
            }
        printf("swap_half %d swap_lbo %d sbo_zero %d : max|D - ref| %.3e  pad/dup columns err %.3e  "
               "16x256b map err %.3e  | cycles issue->commit %lld  issue->regs %lld\n",
               o.swap_half, o.swap_lbo, o.sbo_zero, e32, edup, e16, clk[0], clk[1]);
    }
    printf("\nlatency of the product vs how it is issued (correct packing; cycles at the SM clock)\n");
    const int cfg[][2] = {{1, 16}, {2, 16}, {4, 16}, {8, 16}, {16, 16}, {1, 1}, {1, 2}, {1, 4}, {1, 8}};
    for (auto &c : cfg) {
        Opt o{0, 0, 0, 20, c[0], c[1]};
        probe<<<1, 160>>>(dW, dh, d32, d16, dclk, o);
        cudaDeviceSynchronize();
        long long clk[3];
        cudaMemcpy(clk, dclk, 24, cudaMemcpyDeviceToHost);
        printf("%2d MMA over %2d accumulator(s): issue loop %4lld  issue->commit %4lld  issue->regs %4lld\n", c[1], c[0],
               clk[2], clk[0], clk[1]);
    }
    return 0;
}
