// tc5_probe.cu -- what the recurrence kernel needs to know about tcgen05 on this part,
// measured instead of assumed (tools, not product):
//   P1  packing of a bf16 A operand held in tensor memory (which half of a column is even k)
//   P2  roles of the leading / stride byte offsets for an MN-major, un-swizzled B tile
//       ([unit][8 chunks] rows of 16 bytes -- the layout h_t is exchanged in), and whether
//       a zero stride (second 8-column group = the first) is honoured
//   P3  thread <-> element map of tcgen05.ld.16x256b and a lane base of 32w + 16
//   P4  latency of one recurrent product issued as tcgen05.mma (M = 128, N = 8..64, K = H):
//       first issue -> commit observed -> accumulators in registers, with A in tensor memory
//       (.ts form) and A in shared memory (.ss form, K-major un-swizzled core matrices); the
//       issue loop is fully unrolled with compile-time operands (`latency` kernel below) --
//       the loop of `probe` carries a run-time modulo and measures its own scalar code
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


// ---- P4: the product as the recurrence would issue it --------------------------------
// A[128 gate rows][K = 256] bf16, B = h[K][NC chunks] bf16 (MN-major, rows of NC*2 bytes are
// split in 16-byte pieces: piece p of row k at hs + p * (H*16) + k*16), D[128][NC] fp32.
template <int NC, bool TS, int NMMA, int NACC>
__global__ void __launch_bounds__(160, 1) latency(const float *W, const float *h, float *D, long long *clk,
                                                  int reps, int swap_a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int NP = NC / 8;                        // 16-byte pieces per B row
    unsigned char *hs = smem;                         // NP * H * 16 bytes
    unsigned char *as = smem + 8 * H * 16;            // A, K-major core matrices: [k piece (8 el)][128 rows][16 B]
    uint64_t *bar = reinterpret_cast<uint64_t *>(as + 128 * H * 2);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < H * NC; i += blockDim.x) {
        const int k = i / NC, c = i % NC;
        reinterpret_cast<__nv_bfloat16 *>(hs + (c / 8) * (H * 16) + k * 16)[c % 8] = __float2bfloat16(h[k * 8 + c % 8]);
    }
    for (int i = tid; i < 128 * H; i += blockDim.x) {
        const int r = i / H, k = i % H;
        reinterpret_cast<__nv_bfloat16 *>(as + (k / 8) * 2048 + r * 16)[k % 8] = __float2bfloat16(W[r * H + k]);
    }
    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<512>(smem_u32(tptr));
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tptr;
    if (warp < 4) {
        const float *src = W + (size_t)tid * H;
        for (int kc = 0; kc < H / 16; kc++) {
            uint32_t v[8];
            for (int j = 0; j < 8; j++) {
                __nv_bfloat162 p = __floats2bfloat162_rn(src[kc * 16 + 2 * j], src[kc * 16 + 2 * j + 1]);
                v[j] = *reinterpret_cast<uint32_t *>(&p);
            }
            tmem_st_32x8(tmem + ((uint32_t)(warp * 32) << 16) + kc * 8, v);
        }
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    constexpr uint32_t dcol = 128;
    // B: LBO = K-direction stride of 8-row groups (128 B), SBO = N-direction stride (H*16)
    const uint64_t bdesc = (uint64_t(1) << 46) | (uint64_t((H * 16) >> 4) << 32) | (uint64_t(128 >> 4) << 16) |
                           uint64_t((smem_u32(hs) & 0x3FFFF) >> 4);
    // A (.ss): core matrix = 8 rows x 16 B contiguous; next 8 rows +128 B; next k piece +2048 B
    const uint32_t a_k = 2048, a_m = 128;
    const uint32_t albo = swap_a ? a_m : a_k, asbo = swap_a ? a_k : a_m;
    const uint64_t adesc = (uint64_t(1) << 46) | (uint64_t(asbo >> 4) << 32) | (uint64_t(albo >> 4) << 16) |
                           uint64_t((smem_u32(as) & 0x3FFFF) >> 4);
    const uint32_t idesc = idesc_bf16(128, NC, false, true);
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int rep = 0; rep < reps; rep++) {
        if (tid == 0) {
            t0 = clock64();
#pragma unroll
            for (int k = 0; k < NMMA; k++) {
                const uint32_t d = tmem + dcol + NC * (k % NACC);
                if (TS)
                    mma_ts(d, tmem + (k % 16) * 8, bdesc + (uint64_t)((256 * (k % 16)) >> 4), idesc, k >= NACC);
                else
                    mma_ss(d, adesc + (uint64_t)((4096 * (k % 16)) >> 4), bdesc + (uint64_t)((256 * (k % 16)) >> 4),
                           idesc, k >= NACC);
            }
            t3 = clock64();
            mma_commit(smem_u32(bar));
        }
        mbar_wait(smem_u32(bar), rep & 1);
        fence_after_sync();
        if (tid == 0) t1 = clock64();
        if (warp < 4) {
            uint32_t v[8];
            tmem_ld_32x8(tmem + ((uint32_t)(warp * 32) << 16) + dcol, v);
            tmem_ld_wait();
            if (tid == 0) t2 = clock64();
            if (rep == reps - 1)
                for (int j = 0; j < 8; j++) D[tid * 8 + j] = __uint_as_float(v[j]);
            fence_before_sync();
        }
        __syncthreads();
        fence_after_sync();
        if (tid == 0 && rep == reps - 1) {
            clk[0] = t3 - t0;
            clk[1] = t1 - t0;
            clk[2] = t2 - t0;
        }
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem);
}

template <int NC, bool TS, int NMMA, int NACC>
static void run_latency(const float *dW, const float *dh, float *dD, long long *dclk, const std::vector<float> &ref,
                        int swap_a) {
    const int smem = 8 * H * 16 + 128 * H * 2 + 64;
    cudaFuncSetAttribute(latency<NC, TS, NMMA, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(dD, 0, 128 * 8 * 4);
    latency<NC, TS, NMMA, NACC><<<1, 160, smem>>>(dW, dh, dD, dclk, 20, swap_a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("latency<%d,%d,%d,%d>: CUDA error %s\n", NC, (int)TS, NMMA, NACC, cudaGetErrorString(e)); exit(1); }
    std::vector<float> o(128 * 8);
    long long clk[3];
    cudaMemcpy(o.data(), dD, o.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(clk, dclk, 24, cudaMemcpyDeviceToHost);
    double err = 0;
    if (NMMA == 16 && NACC == 1)
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < 8; c++) err = fmax(err, fabs(o[r * 8 + c] - ref[r * NB + c]));
    printf("N %2d  A in %s%s  %3d MMA over %2d acc: issue loop %4lld  issue->commit %4lld  issue->regs %4lld"
           "  (%.1f cycles/MMA issued)", NC, TS ? "tmem" : "smem", TS ? "" : (swap_a ? " (lbo=m)" : " (lbo=k)"), NMMA, NACC,
           clk[0], clk[1], clk[2], (double)clk[1] / NMMA);
    if (NMMA == 16 && NACC == 1) printf("  max|D - ref| %.2e", err);
    printf("\n");
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
    printf("\nlatency of one recurrent product (M 128, K 256; cycles at the SM clock; unrolled issue)\n");
    float *dD;
    cudaMalloc(&dD, 128 * 8 * 4);
    // (M = 128 needs N % 16 == 0: 8 chunks are padded / duplicated to 16 columns)
    run_latency<16, true, 16, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, true, 16, 4>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, false, 16, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<32, true, 16, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<64, true, 16, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<64, false, 16, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, true, 1, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, true, 4, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, true, 8, 1>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, true, 64, 4>(dW, dh, dD, dclk, ref, 0);
    run_latency<16, false, 64, 4>(dW, dh, dD, dclk, ref, 0);
    run_latency<64, true, 64, 4>(dW, dh, dD, dclk, ref, 0);
    run_latency<64, false, 64, 4>(dW, dh, dD, dclk, ref, 0);
    return 0;
}
