// gemm_tc5.cu -- dense bf16 contractions of the train step on the 5th-generation tensor
// cores: TMA-fed, tcgen05.mma with fp32 accumulators in tensor memory, epilogue fused.
//
// Replaces the library GEMMs behind the reference's nn.LSTM / nn.GRU input projections
// and their input / weight gradients (taiyaki/layers.py:515,633 -> cuDNN), the strided
// convolution (layers.py:795, Conv1d) and the score projection of GlobalNormFlipFlop
// (layers.py:1402-1411: scale * tanh(x W^T + b)).
//
//   C[M x N] (fp32) = A[M x K] * B[N x K]^T          bf16 operands, fp32 accumulate
//
// Either operand may be stored K-major (row = M or N index, K contiguous) or MN-major
// (row = K index, M or N contiguous): the weight-gradient products dW = dY^T X contract
// over the time x batch dimension, which is the slow index of both operands, and the
// tensor core reads them in place through MN-major shared-memory descriptors -- no
// transposed copies.
//
// One CTA per 128 x BN output tile (x k-split), 6 warps:
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of 64 K-elements into a 3-stage ring
//   warp 1      MMA issuer: one thread, 4 tcgen05.mma (K = 16) per stage, tcgen05.commit
//               releases the stage and finally signals the accumulator
//   warps 2-5   epilogue: tcgen05.ld (a warp owns 32 accumulator rows), fused epilogue, then
//               either 128-byte-swizzled staging in the (now idle) operand ring + TMA store,
//               or red.global.add.v4.f32 straight from registers for split-K weight gradients
//               with the unit-major -> gate-major row permutation applied on the way out.
// BN = 256 for wide results (A is then read once per 256 columns: these products are bound by
// L2 -> SM operand traffic), 96 KB of shared memory and 256 tensor-memory columns per CTA: two
// CTAs per SM, so one tile's epilogue (HBM-bound: the fp32 result is 4x the bytes of the
// operands) overlaps the other's main loop.
#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc5.cuh"

namespace ty {
using namespace tc5;

enum { kEpiStore = 0, kEpiBiasTanh = 1, kEpiAtomic = 2, kEpiReduce = 3 };

constexpr int kBM = 128, kBK = 64;

struct GemmArgs {
    int M, N, K;
    int kb_total, kb_per_split;
    const float *bias;
    float scale;
    float *c;           // kEpiAtomic: destination
    int ldc;
    int map_g, map_h;   // kEpiAtomic: destination row = (r % g) * h + r / g when g > 0
    unsigned long long *dbg;   // optional timeline: 8 globaltimer stamps per CTA (tools/gemm_bench.py)
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TY_STAMP(i)                                                                          \
    do {                                                                                     \
        if (g.dbg)                                                                           \
            g.dbg[(size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = gtimer(); \
    } while (0)

template <int BN, int kStages, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
    constexpr uint32_t A_BYTES = kBM * kBK * 2, B_BYTES = BN * kBK * 2, STAGE = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem = (raw + 1023u) & ~1023u;          // 128-byte swizzle atoms are 1024-byte aligned
    const uint32_t bars = smem + kStages * STAGE;
    const uint32_t full0 = bars, empty0 = bars + 8 * kStages, tfull = bars + 16 * kStages;
    const uint32_t tptr = tfull + 8;
    uint32_t *tptr_gen = reinterpret_cast<uint32_t *>(smem_raw + (tptr - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y, ks = blockIdx.z;
    const int kb0 = ks * g.kb_per_split;
    const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);

    if (threadIdx.x == 0) TY_STAMP(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (EPI != kEpiAtomic) tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tptr);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tptr_gen;
    if (threadIdx.x == 0) TY_STAMP(1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t sa = smem + stage * STAGE, sb = sa + A_BYTES, fb = full0 + 8 * stage;
                mbar_expect_tx(fb, STAGE);
                if (A_MN) {
#pragma unroll
                    for (int j = 0; j < kBM / 64; j++)
                        tma_load_2d(sa + j * 8192, &tmA, m_blk * kBM + j * 64, kb * kBK, fb);
                } else {
                    tma_load_2d(sa, &tmA, kb * kBK, m_blk * kBM, fb);
                }
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BN / 64; j++)
                        tma_load_2d(sb + j * 8192, &tmB, n_blk * BN + j * 64, kb * kBK, fb);
                } else {
                    tma_load_2d(sb, &tmB, kb * kBK, n_blk * BN, fb);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(kBM, BN, A_MN, B_MN);
            constexpr uint32_t adv_a = A_MN ? kAdvMNMajor : kAdvKMajor;
            constexpr uint32_t adv_b = B_MN ? kAdvMNMajor : kAdvKMajor;
            int stage = 0, phase = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(full0 + 8 * stage, phase);
                fence_after_sync();
                if (kb == kb0) TY_STAMP(2);
                const uint32_t sa = smem + stage * STAGE, sb = sa + A_BYTES;
                const uint64_t da = A_MN ? desc_mnmajor(sa) : desc_kmajor(sa);
                const uint64_t db = B_MN ? desc_mnmajor(sb) : desc_kmajor(sb);
#pragma unroll
                for (int k = 0; k < kBK / 16; k++)
                    mma_ss(tmem, da + k * adv_a, db + k * adv_b, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                mma_commit(empty0 + 8 * stage);       // stage reusable once these MMAs have read it
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            TY_STAMP(3);
            mma_commit(tfull);                        // accumulator complete
        }
    } else {
        // ===== epilogue (warps 2-5; a warp reads the 32 tensor-memory lanes of its quarter) =====
        const int q = warp & 3;
        const int row0 = m_blk * kBM + q * 32;
        const int n0 = n_blk * BN;
        mbar_wait(tfull, 0);
        fence_after_sync();
        if (warp == 2 && lane == 0) TY_STAMP(4);
        if (row0 < g.M) {
            // every MMA has completed, so the operand ring is free: 4 KB of staging per
            // (warp, 32-column chunk), laid out as the 128-byte swizzle of the C tensor map
            // NBUF staging buffers per warp, reused round-robin when the tile has more chunks
            constexpr int NCHUNK = BN / 32;
            constexpr int RING_BUFS = (kStages * (int)STAGE) / (4 * 4096);
            constexpr int NBUF = NCHUNK < RING_BUFS ? NCHUNK : RING_BUFS;
            const uint32_t stage_base = smem + (uint32_t)q * NBUF * 4096;
#pragma unroll 1
            for (int c = 0; c < NCHUNK; c++) {
                const int col0 = n0 + c * 32;
                if (col0 >= g.N) break;
                uint32_t r[32];
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, r);
                tmem_ld_wait();
                if (EPI == kEpiBiasTanh) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float b = (g.bias && col0 + j < g.N) ? __ldg(g.bias + col0 + j) : 0.0f;
                        r[j] = __float_as_uint(g.scale * tanhf(__uint_as_float(r[j]) + b));
                    }
                }
                if (EPI == kEpiAtomic) {
                    const int row = row0 + lane;
                    if (row < g.M) {
                        const int drow = g.map_g > 0 ? (row % g.map_g) * g.map_h + row / g.map_g : row;
                        float *dst = g.c + (size_t)drow * g.ldc + col0;
                        const bool vec = (g.ldc & 3) == 0 && col0 + 32 <= g.N;
                        if (vec) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                             "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                                             "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                                             : "memory");
                        } else {
                            for (int j = 0; j < 32 && col0 + j < g.N; j++)
                                atomicAdd(dst + j, __uint_as_float(r[j]));
                        }
                    }
                } else {
                    const uint32_t buf = stage_base + (c % NBUF) * 4096;
                    if (NBUF < NCHUNK && c >= NBUF) {      // the store that used this buffer has read it
                        if (lane == 0) tma_wait_group_read<NBUF - 1>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t addr = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * j]),
                                     "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                                     : "memory");
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (EPI == kEpiReduce) tma_reduce_add_2d(&tmC, buf, col0, row0);
                        else tma_store_2d(&tmC, buf, col0, row0);
                        tma_commit_group();
                    }
                }
            }
            // the staging memory must outlive the reads of the bulk stores; the writes themselves
            // complete on their own (kernel boundary)
            if (warp == 2 && lane == 0) TY_STAMP(5);
            if (EPI != kEpiAtomic && lane == 0) tma_wait_group_read<0>();
            if (warp == 2 && lane == 0) TY_STAMP(6);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        fence_after_sync();
        tmem_dealloc<TMEM_COLS>(tmem);
        if (lane == 0) TY_STAMP(7);
    }
}

// ---------------------------------------------------------------------------------
// Persistent form for the result-bound products (projections, input gradients, the strided
// convolution, the score projection): one CTA per SM walks the output tiles; the accumulator
// is double-buffered in tensor memory so the epilogue of tile i (tcgen05.ld -> staging -> TMA
// store, bound by the HBM write of the fp32 result) runs under the main loop of tile i + 1,
// and the operand ring (4 x 48 KB at BN = 256) stays full across tile boundaries -- a TMA load
// takes ~1.5-2.5 us under load, longer than the whole main loop of a K = 256 tile.
//   warp 0  TMA producer      warp 1  MMA issuer      warp 2  tensor-memory allocation
//   warps 4-7  epilogue (warp w reads tensor-memory lanes 32 (w % 4) ..)
template <int BN, int NS, bool B_MN, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tc5_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
    constexpr uint32_t A_BYTES = kBM * kBK * 2, B_BYTES = BN * kBK * 2, STAGE = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    constexpr int NCHUNK = BN / 32;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem = (raw + 1023u) & ~1023u;
    const uint32_t staging = smem + NS * STAGE;                 // 4 warps x 2 x 4 KB
    const uint32_t bars = staging + 4 * 2 * 4096;
    const uint32_t full0 = bars, empty0 = bars + 8 * NS, tfull0 = bars + 16 * NS, tempty0 = tfull0 + 16;
    const uint32_t tptr = tempty0 + 16;
    uint32_t *tptr_gen = reinterpret_cast<uint32_t *>(smem_raw + (tptr - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (g.M + kBM - 1) / kBM, n_blocks = (g.N + BN - 1) / BN;
    const int tiles = m_blocks * n_blocks;
    const int kb_total = g.kb_total;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4);       // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tptr);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tptr_gen;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int n_blk = tile % n_blocks, m_blk = tile / n_blocks;
                for (int kb = 0; kb < kb_total; kb++) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem + stage * STAGE, sb = sa + A_BYTES, fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, STAGE);
                    tma_load_2d(sa, &tmA, kb * kBK, m_blk * kBM, fb);
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; j++)
                            tma_load_2d(sb + j * 8192, &tmB, n_blk * BN + j * 64, kb * kBK, fb);
                    } else {
                        tma_load_2d(sb, &tmB, kb * kBK, n_blk * BN, fb);
                    }
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(kBM, BN, false, B_MN);
            constexpr uint32_t adv_b = B_MN ? kAdvMNMajor : kAdvKMajor;
            int stage = 0, phase = 0, it = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
                const int as = it & 1;
                mbar_wait(tempty0 + 8 * as, ((it >> 1) & 1) ^ 1);     // epilogue has drained this buffer
                fence_after_sync();
                const uint32_t d = tmem + as * BN;
                for (int kb = 0; kb < kb_total; kb++) {
                    mbar_wait(full0 + 8 * stage, phase);
                    fence_after_sync();
                    const uint32_t sa = smem + stage * STAGE, sb = sa + A_BYTES;
                    const uint64_t da = desc_kmajor(sa);
                    const uint64_t db = B_MN ? desc_mnmajor(sb) : desc_kmajor(sb);
#pragma unroll
                    for (int k = 0; k < kBK / 16; k++)
                        mma_ss(d, da + k * kAdvKMajor, db + k * adv_b, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    mma_commit(empty0 + 8 * stage);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                mma_commit(tfull0 + 8 * as);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const uint32_t my_stage = staging + (uint32_t)q * 2 * 4096;
        int it = 0, nstore = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
            const int n_blk = tile % n_blocks, m_blk = tile / n_blocks;
            const int as = it & 1;
            const int row0 = m_blk * kBM + q * 32, n0 = n_blk * BN;
            mbar_wait(tfull0 + 8 * as, (it >> 1) & 1);
            fence_after_sync();
            const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + as * BN;
            int nchunk = (g.N - n0 + 31) / 32;
            nchunk = nchunk > NCHUNK ? NCHUNK : nchunk;
            if (row0 >= g.M) nchunk = 0;
#pragma unroll 1
            for (int c = 0; c < nchunk; c++) {
                const int col0 = n0 + c * 32;
                uint32_t r[32];
                tmem_ld_32x32(tbase + c * 32, r);
                tmem_ld_wait();
                if (c == nchunk - 1) {                 // accumulator drained: the MMA warp may refill it
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * as);
                }
                if (EPI == kEpiBiasTanh) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float b = (g.bias && col0 + j < g.N) ? __ldg(g.bias + col0 + j) : 0.0f;
                        r[j] = __float_as_uint(g.scale * tanhf(__uint_as_float(r[j]) + b));
                    }
                }
                const uint32_t buf = my_stage + (nstore & 1) * 4096;
                if (lane == 0) tma_wait_group_read<1>();     // the store two chunks back has read this buffer
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t addr = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * j]),
                                 "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                                 : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, buf, col0, row0);
                    tma_commit_group();
                }
                nstore++;
            }
            if (nchunk == 0) {                         // rows beyond M: nothing to store, still release
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * as);
            }
        }
        if (lane == 0) tma_wait_group_read<0>();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        fence_after_sync();
        tmem_dealloc<TMEM_COLS>(tmem);
    }
}

// ---- host side ------------------------------------------------------------------
static unsigned long long *g_gemm_dbg = nullptr;      // set by ty_gemm_debug_timeline (tools only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2-D row-major tensor [outer][inner] with a row pitch in bytes; box = {box_inner, box_outer};
// 128-byte swizzle (box_inner * element size must be 128 bytes); out-of-bounds reads give 0.
static bool make_map(CUtensorMap *m, CUtensorMapDataType dt, int esize, const void *ptr, uint64_t inner,
                     uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstr[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    (void)esize;
    return fn(m, dt, 2, const_cast<void *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Ring depth.  Short contractions (the projections: K = 256 .. 1024, result 4x the bytes of the
// operands) are bound by the result write: 2 stages = 64 KB, three CTAs per SM keep enough
// bulk stores in flight.  Long contractions (weight gradients over time x batch, ~45 K-blocks
// per CTA) are bound by operand delivery: 5 stages to cover the TMA latency.
template <int BN, int NS, bool A_MN, bool B_MN, int EPI>
static int launch_gemm_ns(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tc,
                          const GemmArgs &g, int k_splits, cudaStream_t s) {
    constexpr size_t smem = NS * (kBM * kBK * 2 + BN * kBK * 2) + 1024 + 256;
    static_assert(NS * (kBM * kBK * 2 + BN * kBK * 2) >= 4 * 4 * 4096, "staging aliases the ring");
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(gemm_tc5_kernel<BN, NS, A_MN, B_MN, EPI>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
    dim3 grid((g.N + BN - 1) / BN, (g.M + kBM - 1) / kBM, k_splits);
    gemm_tc5_kernel<BN, NS, A_MN, B_MN, EPI><<<grid, 192, smem, s>>>(ta, tb, tc, g);
    return check_launch("gemm_tc5_kernel");
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tc, const GemmArgs &g,
                       int k_splits, cudaStream_t s) {
    return launch_gemm_ns<BN, 3, A_MN, B_MN, EPI>(ta, tb, tc, g, k_splits, s);
}

template <int BN, int NS, bool B_MN, int EPI>
static int launch_persistent(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tc,
                             const GemmArgs &g, cudaStream_t s) {
    constexpr size_t smem = NS * (kBM * kBK * 2 + BN * kBK * 2) + 4 * 2 * 4096 + 1024 + 256;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(gemm_tc5_persistent_kernel<BN, NS, B_MN, EPI>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
    static int sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    const int tiles = ((g.M + kBM - 1) / kBM) * ((g.N + BN - 1) / BN);
    gemm_tc5_persistent_kernel<BN, NS, B_MN, EPI><<<tiles < sms ? tiles : sms, 256, smem, s>>>(ta, tb, tc, g);
    return check_launch("gemm_tc5_persistent_kernel");
}

}  // namespace ty

using namespace ty;

extern "C" void ty_gemm_debug_timeline(void *buf) { g_gemm_dbg = static_cast<unsigned long long *>(buf); }

extern "C" int ty_gemm_bf16(const void *A, int lda, int a_mn, const void *B, int ldb, int b_mn, int M, int N,
                            int K, float *C, int ldc, int epi, const float *bias, float scale, int k_splits,
                            int map_g, int map_h, void *stream) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || epi < 0 || epi > 3) {
        set_error("ty_gemm_bf16: bad argument");
        return TY_EINVAL;
    }
    if ((lda & 7) || (ldb & 7) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) ||
        (epi != kEpiAtomic && (ldc & 3))) {
        set_error("ty_gemm_bf16: operands must be 16-byte aligned with 16-byte row pitch (lda %d ldb %d ldc %d)",
                  lda, ldb, ldc);
        return TY_EINVAL;
    }
    if ((a_mn != 0) != (b_mn != 0) && a_mn) {
        set_error("ty_gemm_bf16: MN-major A with K-major B is not instantiated");
        return TY_EINVAL;
    }
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K;
    g.kb_total = (K + kBK - 1) / kBK;
    if (k_splits < 1) k_splits = 1;
    if (k_splits > 1 && epi != kEpiAtomic && epi != kEpiReduce) {
        set_error("ty_gemm_bf16: k_splits > 1 needs an accumulating epilogue");
        return TY_EINVAL;
    }
    g.kb_per_split = (g.kb_total + k_splits - 1) / k_splits;
    k_splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
    g.bias = bias; g.scale = scale; g.c = C; g.ldc = ldc; g.map_g = map_g; g.map_h = map_h;
    g.dbg = g_gemm_dbg;
    // wide tiles read the A operand once per 256 result columns: the projections are bound by
    // L2 -> SM operand traffic, not by the tensor pipe
    static const bool no_persistent = [] { const char *e = getenv("TY_GEMM_PERSISTENT"); return e && e[0] == '0'; }();
    const bool persistent = !a_mn && k_splits == 1 && (epi == kEpiStore || epi == kEpiBiasTanh) && !no_persistent;
    const int BN = N <= 64 ? 64 : (N <= 128 || !persistent) ? 128 : 256;
    CUtensorMap ta, tb, tc;
    bool ok;
    if (a_mn) ok = make_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, (uint64_t)lda * 2, 64, 64);
    else ok = make_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, (uint64_t)lda * 2, 64, kBM);
    if (b_mn) ok = ok && make_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, N, K, (uint64_t)ldb * 2, 64, 64);
    else ok = ok && make_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, K, N, (uint64_t)ldb * 2, 64, BN);
    if (epi != kEpiAtomic)
        ok = ok && make_map(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, N, M, (uint64_t)ldc * 4, 32, 32);
    else
        tc = ta;
    if (!ok) {
        set_error("ty_gemm_bf16: cuTensorMapEncodeTiled failed (M %d N %d K %d lda %d ldb %d ldc %d)", M, N, K,
                  lda, ldb, ldc);
        return TY_ECUDA;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (persistent) {
#define TY_PERS(BNV, NSV)                                                                          \
    do {                                                                                           \
        if (epi == kEpiStore)                                                                      \
            return b_mn ? launch_persistent<BNV, NSV, true, kEpiStore>(ta, tb, tc, g, s)           \
                        : launch_persistent<BNV, NSV, false, kEpiStore>(ta, tb, tc, g, s);         \
        return b_mn ? launch_persistent<BNV, NSV, true, kEpiBiasTanh>(ta, tb, tc, g, s)            \
                    : launch_persistent<BNV, NSV, false, kEpiBiasTanh>(ta, tb, tc, g, s);          \
    } while (0)
        if (BN == 64) TY_PERS(64, 6);
        if (BN == 128) TY_PERS(128, 5);
        TY_PERS(256, 4);
#undef TY_PERS
    }
#define TY_GEMM(BNV, AM, BM_, EP) return launch_gemm<BNV, AM, BM_, EP>(ta, tb, tc, g, k_splits, s)
#define TY_GEMM_EPI(BNV, AM, BM_)                        \
    switch (epi) {                                       \
        case kEpiStore: TY_GEMM(BNV, AM, BM_, kEpiStore);       \
        case kEpiBiasTanh: TY_GEMM(BNV, AM, BM_, kEpiBiasTanh); \
        case kEpiAtomic: TY_GEMM(BNV, AM, BM_, kEpiAtomic);     \
        default: TY_GEMM(BNV, AM, BM_, kEpiReduce);             \
    }
    if (BN == 64) {
        if (!a_mn && !b_mn) { TY_GEMM_EPI(64, false, false) }
        if (!a_mn && b_mn) { TY_GEMM_EPI(64, false, true) }
        TY_GEMM_EPI(64, true, true)
    }
    if (!a_mn && !b_mn) { TY_GEMM_EPI(128, false, false) }
    if (!a_mn && b_mn) { TY_GEMM_EPI(128, false, true) }
    TY_GEMM_EPI(128, true, true)
#undef TY_GEMM_EPI
#undef TY_GEMM
}
