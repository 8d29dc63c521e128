// batching.cu -- training-batch assembly on the device.
// Replaces, for reads that are resident in HBM, the per-chunk numpy work of
// taiyaki/chunk_selection.py:29-95 (sample_chunks), signal_mapping.py:459-554
// (get_chunk_with_sample_length, get_reference_locations, _get_chunk),
// :680-716 (Chunk.apply_filters) and bin/train_flipflop.py:101-135 (stacking,
// flip-flop coding): with the train step at 6.5 ms that host path (7-90 ms per
// batch of 64 x 4000 samples) is the ceiling.  The host only draws the random
// (read, start sample) candidates; everything that touches signal or labels
// happens here.
//
// Read store (built once): all reads concatenated,
//   dacs     int16 [sum siglen]        dacs_off int64 [R+1]
//   r2s      int32 [sum (reflen+1)]    r2s_off  int64 [R+1]   (Ref_to_signal)
//   ref      int16 [sum reflen]        ref_off  int64 [R+1]   (Reference)
//   lin      float2 [R]  current = dacs * lin.x + lin.y  (offset, range,
//            digitisation, shift and scale of signal_mapping.py:474-477 folded)
// Three launches per batch:
//   chunk_meta_kernel    one warp per candidate: reference span of the window
//                        (the two searchsorted calls), max dwell, the filters
//                        -> reject code, sequence length
//   chunk_select_kernel  one block: the first N accepted candidates in draw
//                        order get slots (what the reference's sampling loop
//                        keeps), label offsets by prefix sum, rejection counts
//   chunk_fill_kernel    signal windows -> indata [T][N] (standardised fp32,
//                        optionally time-reversed), labels -> flip-flop codes
//                        (flipflopfings.py:34-78: flop on odd positions of a
//                        homopolymer run) and, for cat-mod models, the
//                        canonical label / mod category tables
#include "common.cuh"

namespace ty {

enum {               // signal_mapping.py:566-577; order = index into the count array
    kRejPass = 0, kRejEmptySeq = 1, kRejEmptySig = 2, kRejShort = 3, kRejNullMap = 4,
    kRejPathBuffer = 5, kRejMeanDwell = 6, kRejMaxDwell = 7, kRejNum = 8
};

struct BatchArgs {
    const int16_t *dacs; const int64_t *dacs_off;
    const int32_t *r2s; const int64_t *r2s_off;
    const int16_t *ref; const int64_t *ref_off;
    const float2 *lin;
    // candidates
    const int32_t *cand_read;      // [M]
    const int32_t *cand_start;     // [M] first sample of the window; < 0: read too short
    int M, N, T;
    // filters (chunk_selection.py:9-26); use_filters = all four statistics present
    int use_filters;
    float filter_mean_dwell, filter_max_dwell, median_meandwell, mad_meandwell, path_buffer;
    int model_stride;
    // meta
    int32_t *lo, *hi, *code;       // [M]
    // selection
    int32_t *slot_cand;            // [N] candidate of each slot (-1: empty)
    int64_t *seqlen;               // [N]
    int64_t *seqoff;               // [N+1]
    int32_t *counts;               // [kRejNum + 2]: rejection counts, accepted, attempts used
    // fill
    float *indata;                 // [T][N]
    int64_t *seqs, *mod_cats;      // [sum seqlen]
    int reverse, nbase;
    const int32_t *can_labels, *mod_labels;   // cat-mod tables (may be null)
};

__global__ void __launch_bounds__(128) chunk_meta_kernel(const BatchArgs a) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= a.M) return;
    const int start = a.cand_start[c];
    int code = kRejPass, lo = 0, hi = 0;
    if (start < 0) {
        code = kRejShort;
    } else {
        const int r = a.cand_read[c];
        const int32_t *r2s = a.r2s + a.r2s_off[r];
        const int n = (int)(a.r2s_off[r + 1] - a.r2s_off[r]);
        const int end = start + a.T;
        // seq_start = searchsorted(r2s, start, 'right') - 1 ; seq_end = searchsorted(r2s, end, 'left')
        int l = 0, h = n;
        while (l < h) { const int m = (l + h) >> 1; if (r2s[m] <= start) l = m + 1; else h = m; }
        lo = l - 1;
        l = 0; h = n;
        while (l < h) { const int m = (l + h) >> 1; if (r2s[m] < end) l = m + 1; else h = m; }
        hi = l;
        if (hi == lo) {
            code = kRejEmptySeq;
        } else if (a.T == 0) {
            code = kRejEmptySig;
        } else if (a.use_filters) {
            // dwells = diff(r2s[lo:hi]) (signal_mapping.py:534-536): hi - lo - 1 values
            int mx = hi - lo > 1 ? 0 : 1;
            for (int i = lo + lane; i + 1 < hi; i += 32) mx = max(mx, r2s[i + 1] - r2s[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(kFullMask, mx, o));
            const float seq_len = (float)(hi - lo);
            const float mean_dwell = (float)a.T / (seq_len + 0.00000001f);
            if ((float)a.T / (seq_len * (float)a.model_stride) <= a.path_buffer)
                code = kRejPathBuffer;
            else if (fabsf(mean_dwell - a.median_meandwell) > a.filter_mean_dwell * a.mad_meandwell)
                code = kRejMeanDwell;
            else if ((float)mx > a.filter_max_dwell * a.median_meandwell)
                code = kRejMaxDwell;
        }
    }
    if (lane == 0) { a.lo[c] = lo; a.hi[c] = hi; a.code[c] = code; }
}

// One block.  Candidates are consumed in draw order until N have passed, exactly
// like the reference's loop (later candidates do not count as attempts).
__global__ void __launch_bounds__(1024) chunk_select_kernel(const BatchArgs a) {
    __shared__ int s_scan[1024];
    __shared__ int s_cnt[kRejNum + 2];
    const int tid = threadIdx.x;
    if (tid < kRejNum + 2) s_cnt[tid] = 0;
    for (int i = tid; i < a.N; i += blockDim.x) { a.slot_cand[i] = -1; a.seqlen[i] = 0; }
    __syncthreads();
    int base = 0;        // accepted before this tile
    for (int c0 = 0; c0 < a.M; c0 += blockDim.x) {
        const int c = c0 + tid;
        const int code = c < a.M ? a.code[c] : -1;
        const int acc = code == kRejPass;
        s_scan[tid] = acc;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {       // inclusive scan
            const int v = tid >= o ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int before = base + s_scan[tid] - acc;          // accepted among earlier candidates
        if (c < a.M && before < a.N) {                        // still an attempt of the loop
            atomicAdd(&s_cnt[code], 1);
            atomicAdd(&s_cnt[kRejNum + 1], 1);
            if (acc) {
                a.slot_cand[before] = c;
                a.seqlen[before] = a.hi[c] - a.lo[c];
            }
        }
        base += s_scan[blockDim.x - 1];
        __syncthreads();
    }
    if (tid == 0) {
        s_cnt[kRejNum] = min(base, a.N);
        int64_t off = 0;
        for (int i = 0; i < a.N; i++) { a.seqoff[i] = off; off += a.seqlen[i]; }
        a.seqoff[a.N] = off;
    }
    __syncthreads();
    if (tid < kRejNum + 2) a.counts[tid] = s_cnt[tid];
}

// grid (tiles of the window + 1, N): blockIdx.x < tiles copies signal, the last block
// of each slot codes the labels.
constexpr int kFillTile = 1024;

__global__ void __launch_bounds__(256) chunk_fill_kernel(const BatchArgs a) {
    const int n = blockIdx.y;
    const int c = a.slot_cand[n];
    const int tiles = (a.T + kFillTile - 1) / kFillTile;
    if ((int)blockIdx.x < tiles) {
        const int t0 = blockIdx.x * kFillTile;
        if (c < 0) {      // empty slot (fewer than N chunks passed): zeros
            for (int t = t0 + threadIdx.x; t < min(a.T, t0 + kFillTile); t += blockDim.x)
                a.indata[(size_t)t * a.N + n] = 0.f;
            return;
        }
        const int r = a.cand_read[c];
        const int16_t *d = a.dacs + a.dacs_off[r] + a.cand_start[c];
        const float2 lin = a.lin[r];
        for (int t = t0 + threadIdx.x; t < min(a.T, t0 + kFillTile); t += blockDim.x) {
            const int to = a.reverse ? a.T - 1 - t : t;
            a.indata[(size_t)to * a.N + n] = fmaf((float)d[t], lin.x, lin.y);
        }
        return;
    }
    if (c < 0 || threadIdx.x >= 32) return;
    // ---- labels: one warp; position i of the (possibly reversed) sequence ----
    const int lane = threadIdx.x;
    const int r = a.cand_read[c];
    const int lo = a.lo[c], L = a.hi[c] - lo;
    const int16_t *ref = a.ref + a.ref_off[r] + lo;
    int64_t *out = a.seqs + a.seqoff[n];
    int64_t *outm = a.mod_cats ? a.mod_cats + a.seqoff[n] : nullptr;
    int run_start = 0;        // start of the homopolymer run that contains the previous label
    int prev = -1;            // previous (canonical) label
    for (int i0 = 0; i0 < L; i0 += 32) {
        const int i = i0 + lane;
        int lab = -2, mod = 0;
        if (i < L) {
            const int raw = ref[a.reverse ? L - 1 - i : i];
            lab = a.can_labels ? a.can_labels[raw] : raw;
            mod = a.mod_labels ? a.mod_labels[raw] : 0;
        }
        int left = __shfl_up_sync(kFullMask, lab, 1);
        if (lane == 0) left = prev;
        // start of this label's run = latest position <= i where the label changed
        int rs = (i < L && lab != left) ? i : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFullMask, rs, o);
            if (lane >= o) rs = max(rs, v);
        }
        if (rs < 0) rs = run_start;
        if (i < L) {
            out[i] = lab + (((i - rs) & 1) ? a.nbase : 0);      // flop on odd positions of the run
            if (outm) outm[i] = mod;
        }
        run_start = __shfl_sync(kFullMask, rs, 31);
        prev = __shfl_sync(kFullMask, lab, 31);
    }
}

}  // namespace ty

using namespace ty;

extern "C" int ty_batch_counts_len(void) { return kRejNum + 2; }

// All pointers are device pointers.  filters: {filter_mean_dwell, filter_max_dwell,
// median_meandwell, mad_meandwell, path_buffer} or NULL (no filtering), model_stride > 0.
// Outputs: indata [T][N] fp32; seqs / mod_cats [>= N * max seq length] int64 (mod_cats and the
// two label tables may be NULL); seqlen [N] int64; seqoff [N+1] int64; counts
// [ty_batch_counts_len()] int32 = rejection counts in the order pass, emptysequence,
// emptysignal, tooshort, nullmapping, pathbuffer, meandwell, maxdwell, then accepted, attempts.
// scratch: 3 * M + N int32.
extern "C" int ty_sample_chunks(const int16_t *dacs, const int64_t *dacs_off, const int32_t *r2s,
                                const int64_t *r2s_off, const int16_t *ref, const int64_t *ref_off,
                                const float *lin, const int32_t *cand_read,
                                const int32_t *cand_start, int M, int N, int T,
                                const float *filters_host, int model_stride, int reverse, int nbase,
                                const int32_t *can_labels, const int32_t *mod_labels,
                                float *indata, int64_t *seqs, int64_t *mod_cats, int64_t *seqlen,
                                int64_t *seqoff, int32_t *counts, int32_t *scratch, void *stream) {
    if (!dacs || !dacs_off || !r2s || !r2s_off || !ref || !ref_off || !lin || !cand_read ||
        !cand_start || !indata || !seqs || !seqlen || !seqoff || !counts || !scratch || M <= 0 ||
        N <= 0 || T <= 0 || nbase <= 0) {
        set_error("ty_sample_chunks: bad argument");
        return TY_EINVAL;
    }
    BatchArgs a{};
    a.dacs = dacs; a.dacs_off = dacs_off; a.r2s = r2s; a.r2s_off = r2s_off; a.ref = ref;
    a.ref_off = ref_off; a.lin = reinterpret_cast<const float2 *>(lin);
    a.cand_read = cand_read; a.cand_start = cand_start; a.M = M; a.N = N; a.T = T;
    a.use_filters = filters_host != nullptr && model_stride > 0;
    if (a.use_filters) {
        a.filter_mean_dwell = filters_host[0]; a.filter_max_dwell = filters_host[1];
        a.median_meandwell = filters_host[2]; a.mad_meandwell = filters_host[3];
        a.path_buffer = filters_host[4];
    }
    a.model_stride = model_stride;
    a.lo = scratch; a.hi = scratch + M; a.code = scratch + 2 * M; a.slot_cand = scratch + 3 * M;
    a.seqlen = seqlen; a.seqoff = seqoff; a.counts = counts;
    a.indata = indata; a.seqs = seqs; a.mod_cats = mod_cats; a.reverse = reverse; a.nbase = nbase;
    a.can_labels = can_labels; a.mod_labels = mod_labels;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    chunk_meta_kernel<<<(M + 3) / 4, 128, 0, s>>>(a);
    chunk_select_kernel<<<1, 1024, 0, s>>>(a);
    const dim3 grid((T + kFillTile - 1) / kFillTile + 1, N);
    chunk_fill_kernel<<<grid, 256, 0, s>>>(a);
    return check_launch("ty_sample_chunks kernels");
}
