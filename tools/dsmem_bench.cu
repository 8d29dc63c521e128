// dsmem_bench.cu -- how expensive is the per-step all-to-all of the recurrent
// kernels?  A cluster of CL CTAs (128 threads) exchanges 512 B per CTA pair per
// round through distributed shared memory and waits on an mbarrier, with
// different store shapes.  Prints ns per round for each method.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
    uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o;
}
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    const uint32_t a = smem_u32(b);
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq(uint64_t *b, uint32_t parity) {
    const uint32_t a = smem_u32(b);
    asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t ra, uint32_t v, uint32_t rb) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(v), "r"(rb) : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t ra, uint32_t v, uint32_t rb) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %1}, [%2];" ::"r"(ra), "r"(v), "r"(rb) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t ra, uint32_t v, uint32_t rb) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %1, %1, %1}, [%2];" ::"r"(ra), "r"(v), "r"(rb) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t ra, uint32_t v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void remote_arrive_release(uint32_t rb) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t rdst, uint32_t src, uint32_t bytes, uint32_t rb) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rdst), "r"(src), "r"(bytes), "r"(rb) : "memory");
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int HS = 256 + 8;
template <int CL, int M, int NT>
__global__ void __launch_bounds__(NT, 1) xchg(int rounds, int spin, unsigned *sink, long long *cycles) {
    __shared__ __align__(128) unsigned char buf[2][CL * 512 + 1024];
    __shared__ __align__(128) unsigned char stage[512];
    __shared__ __align__(8) uint64_t full[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = ctarank();
    if (tid == 0) {
        // methods 0-4: one local arrive (+ tx bytes); method 5: CL*4 remote arrives
        mbar_init(&full[0], M == 5 ? CL * 4 : 1);
        mbar_init(&full[1], M == 5 ? CL * 4 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    const uint32_t b_base = smem_u32(&buf[0][0]);
    const uint32_t bar_base = smem_u32(&full[0]);
    const uint32_t st_base = smem_u32(&stage[0]);
    uint32_t phase = 0, acc = tid;
    const long long t0 = clock64();
    for (int s = 0; s < rounds; s++) {
        const int cur = s & 1, nxt = cur ^ 1;
        if (M != 5 && tid == 0 && s + 1 < rounds) expect_tx(&full[nxt], CL * (M == 10 ? 256 : M == 11 ? 128 : 512));
        if (s > 0) {
            if (M == 5) mbar_wait_acq(&full[cur], (phase >> cur) & 1u); else mbar_wait(&full[cur], (phase >> cur) & 1u);
            phase ^= 1u << cur;
            acc += *reinterpret_cast<volatile uint32_t *>(&buf[cur][(tid * 4) % (CL * 512)]);
        }
        for (int i = 0; i < spin; i++) acc = acc * 1664525u + 1013904223u;   // stand-in for the step's math
        if (s + 1 < rounds) {
            const uint32_t dst0 = b_base + nxt * (CL * 512 + 1024);
            const uint32_t bar = bar_base + nxt * 8;
            if (M == 0) {
                // like the LSTM kernel: lane (r = lane>>2, q = lane&3): row 2q+(r&1), 4-byte pair r>>1 of warp's 16 B
                const int r = lane >> 2, q = lane & 3;
                const uint32_t off = (uint32_t)((2 * q + (r & 1)) * (CL * 64 + 16) + (rank * 64 + warp * 16 + (r >> 1) * 4));
                for (uint32_t p = 0; p < CL; p++) st_async_b32(mapa(dst0 + off, p), acc, mapa(bar, p));
            } else if (M == 1) {
                // b32, 128 B contiguous per warp instruction
                const uint32_t off = rank * 512 + warp * 128 + lane * 4;
                for (uint32_t p = 0; p < CL; p++) st_async_b32(mapa(dst0 + off, p), acc, mapa(bar, p));
            } else if (M == 10) {
                // half the bytes: only lanes with q < 2 send (4 chunks per cluster)
                const uint32_t off = rank * 512 + warp * 128 + lane * 4;
                if ((lane & 3) < 2)
                    for (uint32_t p = 0; p < CL; p++) st_async_b32(mapa(dst0 + off, p), acc, mapa(bar, p));
            } else if (M == 11) {
                // a quarter of the bytes: one warp sends
                const uint32_t off = rank * 512 + lane * 4;
                if (warp == 0)
                    for (uint32_t p = 0; p < CL; p++) st_async_b32(mapa(dst0 + off, p), acc, mapa(bar, p));
            } else if (M == 2) {
                // v2: 256 B contiguous per warp instruction; warp w covers half-blocks for peers
                const uint32_t off = rank * 512 + (warp & 1) * 256 + lane * 8;
                for (uint32_t p = 0; p < CL / 2; p++) {
                    const uint32_t peer = (warp >> 1) * (CL / 2) + p;
                    st_async_v2(mapa(dst0 + off, peer), acc, mapa(bar, peer));
                }
            } else if (M == 3) {
                // v4: 512 B contiguous per warp instruction; warp w serves peers w*CL/4 ..
                const uint32_t off = rank * 512 + lane * 16;
                for (uint32_t p = 0; p < CL / 4; p++) {
                    const uint32_t peer = warp * (CL / 4) + p;
                    st_async_v4(mapa(dst0 + off, peer), acc, mapa(bar, peer));
                }
            } else if (M == 6) {
                // staged v4: every thread writes 4 B to local staging, CTA barrier, then as M3 from smem
                *reinterpret_cast<uint32_t *>(&stage[tid * 4]) = acc;
                __syncthreads();
                const uint4 v = *reinterpret_cast<const uint4 *>(&stage[lane * 16]);
                const uint32_t off = rank * 512 + lane * 16;
                for (uint32_t p = 0; p < CL / 4; p++) {
                    const uint32_t peer = warp * (CL / 4) + p;
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                 ::"r"(mapa(dst0 + off, peer)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mapa(bar, peer)) : "memory");
                }
            } else if (M == 4) {
                // bulk copy: stage locally, CTA barrier, CL lanes of warp 0 each copy 512 B
                *reinterpret_cast<uint32_t *>(&stage[tid * 4]) = acc;
                fence_async();
                __syncthreads();
                if (tid < CL) bulk_copy(mapa(dst0 + rank * 512, tid), st_base, 512, mapa(bar, tid));
            } else if (M == 8) {
                // b32, the LSTM kernel's lane->(row, pair) permutation inside one 128 B block per warp
                const int r = lane >> 2, q = lane & 3;
                const uint32_t off = rank * 512 + warp * 128 + (2 * q + (r & 1)) * 16 + (r >> 1) * 4;
                for (uint32_t p = 0; p < CL; p++) st_async_b32(mapa(dst0 + off, p), acc, mapa(bar, p));
            } else if (M == 7) {
                // 8 warps, one cell per thread: warp w = 2*wp + h fills bytes h*8.. of the 16 B rows of
                // block wp; 16 lanes -> peer p, 16 lanes -> peer p + CL/2
                const int g8 = lane >> 2, q = lane & 3;
                const int wp = warp >> 1, h = warp & 1;
                const uint32_t off = rank * 512 + wp * 128 + (2 * q + (g8 & 1)) * 16 + h * 8 + (g8 >> 2) * 4;
                const uint32_t peer0 = ((g8 >> 1) & 1) ? CL / 2 : 0;
                for (uint32_t p = 0; p < CL / 2; p++) st_async_b32(mapa(dst0 + off, peer0 + p), acc, mapa(bar, peer0 + p));
            } else if (M == 9) {
                // 8 warps, one cell per thread, warp-private 64 B block (rows of 8 B)
                const int g8 = lane >> 2, q = lane & 3;
                const uint32_t off = rank * 512 + warp * 64 + (2 * q + (g8 & 1)) * 8 + (g8 >> 2) * 4;
                const uint32_t peer0 = ((g8 >> 1) & 1) ? CL / 2 : 0;
                for (uint32_t p = 0; p < CL / 2; p++) st_async_b32(mapa(dst0 + off, peer0 + p), acc, mapa(bar, peer0 + p));
            } else if (M == 5) {
                // plain remote stores + one release-arrive per warp per peer
                const uint32_t off = rank * 512 + warp * 128 + (lane & 7) * 16;
                const uint32_t peer0 = (lane >> 3) * (CL / 4);
                for (uint32_t p = 0; p < CL / 4; p++) st_cluster_v4(mapa(dst0 + off, peer0 + p), acc);
                __syncwarp();
                if (lane < CL) remote_arrive_release(mapa(bar, lane));
            }
        }
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
    cluster_sync_all();
}

template <int CL, int M, int NT = 128>
static int run(const char *name, int clusters, int rounds, int spin) {
    unsigned *sink; long long *cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, 8));
    auto k = xchg<CL, M, NT>;
    if (CL > 8) CK(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(NT);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        cudaEventRecord(a);
        CK(cudaLaunchKernelEx(&cfg, k, rounds, spin, sink, cyc));
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("CL=%2d clusters=%d spin=%4d %-34s %8.1f ns/round  %7.1f cycles/round\n", CL, clusters, spin, name,
           best * 1e6 / rounds, (double)c / rounds);
    cudaFree(sink); cudaFree(cyc);
    return 0;
}

// Two-phase reception: every sender serves its peers in rotated order (rank+1, rank+2, ...,
// itself last), the first four through barrier A of the receiver, the last four through
// barrier B.  If the sender's egress is what spreads the arrivals, A completes early and the
// receiver can start on that half while B is still in flight.  TWO = false: same rotated
// order, one barrier, all the work after it.
template <int CL, bool TWO>
__global__ void __launch_bounds__(128, 1) xchg2(int rounds, int work, unsigned *sink, long long *cycles) {
    __shared__ __align__(128) unsigned char buf[2][CL * 512];
    __shared__ __align__(8) uint64_t fullA[2], fullB[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = ctarank();
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&fullA[i], 1); mbar_init(&fullB[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    const uint32_t b_base = smem_u32(&buf[0][0]);
    const uint32_t barA = smem_u32(&fullA[0]), barB = smem_u32(&fullB[0]);
    uint32_t phase = 0, acc = tid;
    const long long t0 = clock64();
    for (int s = 0; s < rounds; s++) {
        const int cur = s & 1, nxt = cur ^ 1;
        if (tid == 0 && s + 1 < rounds) {
            if (TWO) { expect_tx(&fullA[nxt], CL / 2 * 512); expect_tx(&fullB[nxt], CL / 2 * 512); }
            else expect_tx(&fullA[nxt], CL * 512);
        }
        if (s > 0) {
            mbar_wait(&fullA[cur], (phase >> cur) & 1u);
            acc += *reinterpret_cast<volatile uint32_t *>(&buf[cur][(tid * 4) % (CL * 512)]);
            if (TWO) {
                for (int i = 0; i < work / 2; i++) acc = acc * 1664525u + 1013904223u;
                mbar_wait(&fullB[cur], (phase >> cur) & 1u);
                acc += *reinterpret_cast<volatile uint32_t *>(&buf[cur][(tid * 4 + 2048) % (CL * 512)]);
                for (int i = 0; i < work / 2; i++) acc = acc * 1664525u + 1013904223u;
            } else {
                for (int i = 0; i < work; i++) acc = acc * 1664525u + 1013904223u;
            }
            phase ^= 1u << cur;
        }
        if (s + 1 < rounds) {
            const uint32_t off = rank * 512 + warp * 128 + lane * 4;
            const uint32_t dst0 = b_base + nxt * (CL * 512) + off;
#pragma unroll
            for (uint32_t p = 0; p < CL; p++) {
                const uint32_t peer = (rank + 1 + p) & (CL - 1);
                const uint32_t bar = ((TWO && p >= CL / 2) ? barB : barA) + nxt * 8;
                st_async_b32(mapa(dst0, peer), acc, mapa(bar, peer));
            }
        }
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
    cluster_sync_all();
}

template <int CL, bool TWO>
static int run2(const char *name, int clusters, int rounds, int work) {
    unsigned *sink; long long *cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, 8));
    auto k = xchg2<CL, TWO>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(128);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int it = 0; it < 3; it++) { CK(cudaLaunchKernelEx(&cfg, k, rounds, work, sink, cyc)); CK(cudaDeviceSynchronize()); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("CL=%2d clusters=%d work=%4d %-40s %7.1f cycles/round\n", CL, clusters, work, name, (double)c / rounds);
    cudaFree(sink); cudaFree(cyc);
    return 0;
}

int main() {
    const int R = 4000;
    for (int work : {0, 60, 120}) {
        run2<8, false>("rotated order, one barrier", 8, R, work);
        run2<8, true>("rotated order, two-phase reception", 8, R, work);
    }
    for (int spin : {0}) {
        for (int clusters : {8}) {
            run<8, 0>("b32 rows (as LSTM kernel)", clusters, R, spin);
            run<8, 1>("b32 128B-contiguous", clusters, R, spin);
            run<8, 2>("v2.b32 256B-contiguous", clusters, R, spin);
            run<8, 3>("v4.b32 512B-contiguous", clusters, R, spin);
            run<8, 8>("b32 128B block, LSTM lane order", clusters, R, spin);
            run<8, 7, 256>("8 warps, interleaved 8B pieces", clusters, R, spin);
            run<8, 9, 256>("8 warps, private 64B blocks", clusters, R, spin);
            run<8, 6>("staged + syncthreads + v4", clusters, R, spin);
            run<8, 4>("staged + bulk copy 512B", clusters, R, spin);
            run<8, 5>("st.cluster.v4 + release arrive", clusters, R, spin);
        }
    }
    run<8, 10>("b32, 256 B per pair", 8, R, 0);
    run<8, 10>("b32, 256 B per pair, 16 clusters", 16, R, 0);
    run<8, 11>("b32, 128 B per pair", 8, R, 0);
    run<8, 1>("b32 128B-contiguous, 16 clusters", 16, R, 0);
    run<4, 1>("CL=4 b32 512 B per pair", 8, R, 0);
    run<2, 1>("CL=2 b32 512 B per pair", 8, R, 0);
    run<4, 3>("CL=4 v4 512 B per pair", 8, R, 0);
    run<2, 11>("CL=2 b32 128 B per pair", 8, R, 0);
    run<16, 0>("b32 rows", 8, R, 0);
    run<16, 3>("v4.b32 512B-contiguous", 8, R, 0);
    run<16, 4>("staged + bulk copy 512B", 8, R, 0);
    run<16, 8>("b32 128B block, LSTM lane order", 8, R, 0);
    return 0;
}
