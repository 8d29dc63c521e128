/*
 * taiyaki_b200.h -- C ABI of libtaiyaki_b200.so, the B200 (sm_100a) drop-in
 * for the flip-flop CRF training hot path of nanoporetech/taiyaki v5.3.0.
 *
 * Two layers are exported:
 *
 *  (A) HOST-POINTER DROP-INS with the reference's exact names and signatures
 *      (taiyaki/ctc/libctc.pxd:3-25, c_crf_flipflop.h, c_cat_mod_flipflop.h).
 *      Linking taiyaki/ctc/ctc.pyx against this library instead of
 *      c_crf_flipflop.c / c_cat_mod_flipflop.c needs no source change; the
 *      copies to and from the device happen inside the call.
 *
 *  (B) DEVICE-POINTER entry points (`ty_*`), asynchronous on a caller-supplied
 *      CUDA stream with caller-supplied workspace, used by the Python operator
 *      layer (taiyaki_b200/ctc.py, layers.py) so that scores never leave HBM.
 *      They replace, per function:
 *        ty_crf_flipflop_*      c_crf_flipflop.c:255-290 (cost), :434-516 (grad)
 *        ty_cat_mod_flipflop_*  c_cat_mod_flipflop.c:286-345, :493-582
 *        ty_flipflop_indices    ctc.pyx:127-132, :287-292 (flipflopfings.py:6-31)
 *        ty_flipflop_logz*      cupy_extensions/flipflop.py:10-368 and
 *                               layers.py:1253-1299 (logZ and its gradient)
 *        ty_lstm_* / ty_gru_*   the cuDNN calls behind layers.py:515 (nn.LSTM)
 *                               and layers.py:633 (nn.GRU), bias_hh == 0
 *
 * All tensors are C-contiguous fp32 unless noted; score tensors are
 * [nblk][nbatch][ntrans] exactly as the reference lays them out.  No function
 * allocates device memory except the (A) layer, which keeps a grow-only pool.
 *
 * Return value of every ty_* function: 0 on success, else a TY_E* code.
 */
#ifndef TAIYAKI_B200_H
#define TAIYAKI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TY_OK 0
#define TY_EINVAL 1   /* bad shape / null pointer / unsupported size      */
#define TY_EWORKSPACE 2 /* workspace too small                            */
#define TY_ECUDA 3    /* CUDA runtime error (see ty_last_error_string)    */

const char *ty_last_error_string(void);
const char *ty_version(void);

/* ---------------------------------------------------------------- (A) ---
 * Reference ABI, host pointers.  Index arrays are size_t and packed as the
 * reference packs them: stayidxs has sum(seqlen) entries, moveidxs (and the
 * mod arrays) sum(seqlen) - nbatch, chunk b starting at sum(seqlen[:b]) (- b).
 * score[b] = 0.5 * (forward + backward) log-score; grad rows sum to one.
 */
void crf_flipflop_grad(const float *logprob, size_t ntrans, size_t nblk,
                       size_t nbatch, const size_t *moveidxs,
                       const size_t *stayidxs, const int32_t *seqlen,
                       float *score, float *grad);
void crf_flipflop_cost(const float *logprob, size_t ntrans, size_t nblk,
                       size_t nbatch, const size_t *moveidxs,
                       const size_t *stayidxs, const int32_t *seqlen,
                       float *score);
void cat_mod_flipflop_grad(const float *logprob, size_t ntrans, size_t nblk,
                           size_t nbatch, const size_t *moveidxs,
                           const size_t *stayidxs, const size_t *modmoveidxs,
                           const float *modmovefacts, const int32_t *seqlen,
                           float *score, float *grad);
void cat_mod_flipflop_cost(const float *logprob, size_t ntrans, size_t nblk,
                           size_t nbatch, const size_t *moveidxs,
                           const size_t *stayidxs, const size_t *modmoveidxs,
                           const float *modmovefacts, const int32_t *seqlen,
                           float *score);

/* Tuning knobs for A/B timing (tools/microbench.py); negative = leave unchanged.
 * forced_p: DP positions per thread (1, 2, 4, 8, 16; 0 = automatic, default, or TY_CRF_P);
 * fused: 1 = chains with the posterior fused in where the chunk fits (default), 0 = always
 * the chain kernel + posterior kernel pair (or TY_CRF_FUSED=0). */
void ty_crf_tuning(int forced_p, int fused);
/* Which kernels the last ty_crf_flipflop call of this process launched: 1 = chain kernel +
 * posterior kernel, 2 = chains with the posterior fused in, 3 = cost only (0 = none yet). */
int ty_crf_last_path(void);

/* ---------------------------------------------------------------- (B) ---
 * Label-constrained CRF, device pointers.
 *
 * moveidx/stayidx/modmoveidx are int32 in the reference packing.  seqlen is
 * int32 [nbatch] ON THE DEVICE; max_seqlen (host value, >= every seqlen) only
 * selects the launch configuration.  `sharp` multiplies columns < nsharp
 * before use (ctc.pyx:119 and :265-269); pass nsharp = ntrans for the plain
 * model.  score_out[b] = score_scale * score, grad_out = grad_scale * G, so
 * the operator's -x/nblk (ctc.pyx:66,113) costs nothing extra.
 * modmoveidx == NULL selects the plain model.  grad_out == NULL computes the
 * forward score only (crf_flipflop_cost semantics: forward score, not the
 * forward/backward average).
 */
size_t ty_crf_flipflop_workspace_bytes(int ntrans, int nblk, int nbatch,
                                       int max_seqlen, int want_grad);

int ty_crf_flipflop(const float *logprob, int ntrans, int nblk, int nbatch,
                    const int32_t *moveidx, const int32_t *stayidx,
                    const int32_t *modmoveidx, const float *modmovefact,
                    const int32_t *seqlen, int max_seqlen,
                    float sharp, int nsharp,
                    float score_scale, float *score_out,
                    float grad_scale, float *grad_out,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Flip-flop index build on the device.  seqs: int64 [total] concatenated
 * flip-flop coded labels; seqlen int64 [nbatch]; outputs int32 in the
 * reference packing plus seqlen32.  mod_cats (int64 [total]) / can_mods_offsets
 * (int32 [nbase+1]) / mod_cat_weights (float [ncan+nmod]) may be NULL. */
int ty_flipflop_indices(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                        int64_t total, int nbase,
                        const int64_t *mod_cats, const int32_t *can_mods_offsets,
                        const float *mod_cat_weights,
                        int32_t *moveidx, int32_t *stayidx, int32_t *seqlen32,
                        int32_t *modmoveidx, float *modmovefact, void *stream);

/* As ty_flipflop_indices; additionally ORs 1 into *bad_flag (device int32, may be
 * NULL) when a label lies outside [0, 2 nbase) or a modification category outside its
 * base's range -- the host assertions of ctc.pyx:133-134 as a flag the caller reads
 * back whenever it next copies something to the host.  Bad values are clamped, so the
 * index arrays are always safe to use. */
int ty_flipflop_indices_checked(const int64_t *seqs, const int64_t *seqlen, int nbatch,
                                int64_t total, int nbase, const int64_t *mod_cats,
                                const int32_t *can_mods_offsets, const float *mod_cat_weights,
                                int32_t *moveidx, int32_t *stayidx, int32_t *seqlen32,
                                int32_t *modmoveidx, float *modmovefact, int32_t *bad_flag,
                                void *stream);

/* Partition function over the 2*nbase-state lattice.
 * scores: [nblk][nbatch] rows of `ld` floats of which the first
 * S = 2*nbase*(nbase+1) are transition scores (ld >= S lets the cat-mod
 * tensor be used in place, train_flipflop.py:175).
 * logz_out[b] = logz_scale * logZ_b.  If grad_out != NULL it receives
 * grad_scale * d logZ / d scores (rows of ld_grad floats; columns >= S are
 * left untouched), optionally accumulated into what is already there. */
size_t ty_flipflop_logz_workspace_bytes(int nbase, int nblk, int nbatch);

int ty_flipflop_logz(const float *scores, int ld, int nblk, int nbatch,
                     int nbase, float logz_scale, float *logz_out,
                     float grad_scale, float *grad_out, int ld_grad,
                     int accumulate, void *workspace, size_t workspace_bytes,
                     void *stream);

/* The same in two phases (bit 0: lattice chains, bit 1: posterior), so the
 * chains can be overlapped with other work. */
int ty_flipflop_logz_phase(const float *scores, int ld, int nblk, int nbatch,
                           int nbase, float logz_scale, float *logz_out,
                           float grad_scale, float *grad_out, int ld_grad,
                           int accumulate, void *workspace, size_t workspace_bytes,
                           int phases, void *stream);

/* The whole training loss of bin/train_flipflop.py:163-182 in one call:
 *   cost_out[b] = CRF cost (= -score / nblk / sharp), logz_out[b] = logZ_b / nblk,
 *   grad_out    = d (cost_b + logZ_b / nblk) / d scores   [nblk][nbatch][ntrans]
 * The two families of chains run concurrently (side stream inside the library);
 * ncan = number of stay/move transition columns (40), the remaining
 * ntrans - ncan columns are the cat-mod stream (modmoveidx != NULL). */
size_t ty_flipflop_train_loss_workspace_bytes(int ntrans, int nblk, int nbatch,
                                              int max_seqlen, int want_grad);
int ty_flipflop_train_loss(const float *scores, int ntrans, int nblk, int nbatch,
                           const int32_t *moveidx, const int32_t *stayidx,
                           const int32_t *modmoveidx, const float *modmovefact,
                           const int32_t *seqlen, int max_seqlen, float sharp,
                           int ncan, float *cost_out, float *logz_out,
                           float *grad_out, void *workspace, size_t workspace_bytes,
                           void *stream);

/* ----------------------------------------------------------------------
 * Recurrent layers (time-major [T][N][H], PyTorch gate order, b_hh == 0).
 * See taiyaki_b200/csrc/rnn_fp32.cu / rnn_common.cuh for the data layout of `reserve`.
 * xproj: [T][N][G*H] = x W_ih^T computed by the caller (one large GEMM);
 * bias: [G*H] input bias b_ih added inside the kernel (NULL if already in
 * xproj); w_hh: [G*H][H] fp32; reverse != 0 iterates t downward
 * (layers.py:117-153 without the two flips).  G = 4 (LSTM) or 3 (GRU).
 * Backward writes dxproj [T][N][G*H] (gradient of xproj), for the GRU also
 * dhn [T][N][H] (gradient of the hidden-side n pre-activation W_hn h, which
 * differs from the x-side one by the reset gate), and ADDS the bias gradient
 * (sum of dxproj over time and chunks) into dbias [G*H] when it is not NULL.
 * Weight and input gradients are dense GEMMs the caller does over all steps.
 */
size_t ty_rnn_reserve_bytes(int cell, int T, int N, int H);

int ty_lstm_forward(const float *xproj, const float *bias, const float *w_hh,
                    int T, int N, int H, int reverse, float *y, void *reserve,
                    void *stream);
int ty_lstm_backward(const float *dy, const float *w_hh, int T, int N, int H,
                     int reverse, const float *y, const void *reserve,
                     float *dxproj, float *dbias, void *stream);
int ty_gru_forward(const float *xproj, const float *bias, const float *w_hh,
                   int T, int N, int H, int reverse, float *y, void *reserve,
                   void *stream);
int ty_gru_backward(const float *dy, const float *w_hh, int T, int N, int H,
                    int reverse, const float *y, const void *reserve,
                    float *dxproj, float *dhn, float *dbias, void *stream);

/* Extended forms used by the Python layer (cell: 0 = LSTM, 1 = GRU).
 * y_bf16 (may be NULL): additionally write y as bf16 [T][N][H], the operand of
 * the next layer's projection GEMM and of this layer's weight-gradient GEMM.
 * grads_bf16 != 0: dxproj / dhn point to bf16 buffers (same shapes); they are
 * only ever GEMM operands, and the bias gradient is accumulated in fp32 before
 * rounding. */
int ty_rnn_forward_ex(int cell, const float *xproj, const float *bias,
                      const float *w_hh, int T, int N, int H, int reverse,
                      float *y, void *y_bf16, void *reserve, void *stream);
int ty_rnn_backward_ex(int cell, const float *dy, const float *w_hh, int T, int N,
                       int H, int reverse, const float *y, const void *reserve,
                       void *dxproj, void *dhn, int grads_bf16, float *dbias,
                       void *stream);

/* UNIT-MAJOR forms (taiyaki_b200/csrc/rnn_ws.cu: warp-specialised kernels with
 * a dedicated I/O warp).  Same semantics as the _ex forms, different layout of
 * the projection tensors: the G gates of a (chunk, unit) cell are adjacent,
 *     xproj       [T][N][H][G] fp32    = x (P W_ih)^T,  P = gate-major -> unit-major
 *                                        row permutation of W_ih (row u*G+g <- g*H+u)
 *     dxproj_bf16 [T][N][H][G] bf16    gradient of xproj
 *     dhid_bf16   [T][N][H][3] bf16    GRU only: (dr, dz, d(W_hn h)), the gradient
 *                                        of the hidden-side product h W_hh^T
 * so a CTA's share of a chunk is one contiguous run.  bias, w_hh and dbias stay
 * gate-major ([G*H]), i.e. they are the parameters themselves; `reserve` has
 * ty_rnn_reserve_bytes() bytes and is private to the forward / backward pair.
 * The weight gradients the caller forms from dxproj / dhid come out with the
 * same row permutation P. */
/* 1 when the warp-specialised kernels cover this hidden size (multiples of 64 up to 448
 * with clusters of 8; 32, 96, 160, 224 with clusters of 4 -- the reference's defaults are
 * 256 / 384 and 96 for its "fast" models, bin/_bin_argparse.py:16, README.md:354-359);
 * other sizes run through the gate-major fp32 entry points. */
int ty_rnn_um_supported(int hidden);
int ty_rnn_forward_um(int cell, const float *xproj, const float *bias,
                      const float *w_hh, int T, int N, int H, int reverse,
                      float *y, void *y_bf16, void *reserve, void *stream);
int ty_rnn_backward_um(int cell, const float *dy, const float *w_hh, int T, int N,
                       int H, int reverse, const float *y, const void *reserve,
                       void *dxproj_bf16, void *dhid_bf16, float *dbias,
                       void *stream);

/* Scatter half of the time-major 1-D convolution backward (layers.py:744-850):
 * dx[T][N][C] from the column gradient dcols[T_out][N][C*k] (window j of output
 * step to covers input sample to*stride + j - pad_left). */
int ty_col2im_time_major(const float *dcols, int Tout, int N, int C, int k,
                         int stride, int pad_left, int T, float *dx, void *stream);

/* Same with a row stride ld >= C*k of dcols (GEMM operands padded to 8 columns). */
int ty_col2im_time_major_ld(const float *dcols, int ld, int Tout, int N, int C, int k,
                            int stride, int pad_left, int T, float *dx, void *stream);

/* Window gather of the convolution (layers.py:816-831) straight into the GEMM
 * operand: cols[to*N + n][c*k + j] = x[to*stride + j - pad_left][n][c] as bf16,
 * [Tout*N][ld] with ld a multiple of 8 and > C*k; column C*k is 1.0 (so the bias
 * rides in the GEMM as column C*k of the weight operand and its gradient falls
 * out of the weight-gradient GEMM), further columns 0. */
int ty_im2col_time_major_bf16(const float *x, int T, int N, int C, int k, int stride,
                              int pad_left, int Tout, int ld, void *cols_bf16,
                              void *stream);

/* Direct small convolutions (stride 1; the 1->4 and 4->16 channel, window 5
 * layers of models/mLstm_flipflop.py), time-major fp32 [T][N][C].
 * act: 0 linear, 1 tanh, 2 swish (taiyaki/activation.py).  forward writes the
 * pre-activation z and a = act(z), both [T][N][Cout]; backward takes da, uses
 * dz [T][N][Cout] as scratch, ACCUMULATES into dW [Cout][C][k] and db [Cout]
 * (db may be NULL) and writes dx [T][N][C] unless it is NULL.
 * ty_conv_small_supported() tells which (C, Cout, k) are instantiated. */
int ty_conv_small_supported(int C, int Cout, int k);
/* Single-channel strided feature convolution (first layer of the reference's mGru models,
 * Convolution(1, size, 19, stride=2); layers.py:795) as direct fp32 kernels: x [T][N][1],
 * w [Cout][k], z [Tout][N][Cout] = b + conv (pre-activation); ty_conv_in1_wgrad ADDS the weight
 * and bias gradients of dz into dw [Cout][k] / db [Cout] (zero them first).  The input is the
 * signal, so there is no input gradient.  Supported: C == 1, k <= 32. */
int ty_conv_in1_supported(int C, int Cout, int k);
int ty_conv_in1_forward(const float *x, const float *w, const float *b, int T, int N, int Cout,
                        int k, int stride, int pad_left, int Tout, float *z, void *stream);
int ty_conv_in1_wgrad(const float *x, const float *dz, int T, int N, int Cout, int k, int stride,
                      int pad_left, int Tout, float *dw, float *db, void *stream);
int ty_conv_small_forward(const float *x, const float *w, const float *b, int T, int N,
                          int C, int Cout, int k, int pad_left, int act, float *z,
                          float *a, void *stream);
int ty_conv_small_backward(const float *da, const float *z, const float *x,
                           const float *w, int T, int N, int C, int Cout, int k,
                           int pad_left, int act, float *dz, float *dW, float *db,
                           float *dx, void *stream);

/* ----------------------------------------------------------------------
 * Training-batch assembly on the device (taiyaki/chunk_selection.py:29-95,
 * signal_mapping.py:459-554 and :680-716, bin/train_flipflop.py:101-135) for
 * reads resident in HBM: concatenated DACs (int16), Ref_to_signal (int32,
 * reflen + 1 entries per read) and Reference (int16) with int64 offset arrays
 * of R + 1 entries, lin [R][2] with current = dacs * lin[r][0] + lin[r][1].
 * The caller draws M candidate (read, first sample) pairs in attempt order
 * (first sample < 0: read shorter than the chunk); the first N that pass the
 * filters fill the batch.  filters_host: HOST pointer to {filter_mean_dwell,
 * filter_max_dwell, median_meandwell, mad_meandwell, path_buffer} or NULL.
 * Outputs (device): indata [T][N] fp32 (time reversed if reverse != 0), seqs
 * (flip-flop coded, concatenated in slot order) and optionally mod_cats (with
 * the two label tables of a cat-mod model), seqlen [N], seqoff [N + 1], counts
 * [ty_batch_counts_len()] = per-reason attempt counts (pass, emptysequence,
 * emptysignal, tooshort, nullmapping, pathbuffer, meandwell, maxdwell),
 * accepted, attempts.  scratch: 3 M + N int32. */
int ty_batch_counts_len(void);
int ty_sample_chunks(const int16_t *dacs, const int64_t *dacs_off, const int32_t *r2s,
                     const int64_t *r2s_off, const int16_t *ref, const int64_t *ref_off,
                     const float *lin, const int32_t *cand_read,
                     const int32_t *cand_start, int M, int N, int T,
                     const float *filters_host, int model_stride, int reverse, int nbase,
                     const int32_t *can_labels, const int32_t *mod_labels,
                     float *indata, int64_t *seqs, int64_t *mod_cats, int64_t *seqlen,
                     int64_t *seqoff, int32_t *counts, int32_t *scratch, void *stream);

/* Best path over the flip-flop lattice (taiyaki/decode.py:79-115,
 * cupy_extensions/flipflop.py:387-518): scores [T][N][S] fp32 -> fwd
 * [T+1][N][2 nbase] max-scores (flip states start at 0, flop states at -1e30),
 * traceback [T][N][2 nbase] int64 (best predecessor state, lowest index on a
 * tie), path [T+1][N] int64.  nbase == 4. */
int ty_flipflop_viterbi(const float *scores, int T, int N, int nbase, float *fwd,
                        int64_t *traceback, int64_t *path, void *stream);

/* Best alignment of label sequences to flip-flop transition scores -- the max-product
 * twin of the training DP (taiyaki/flipflop_remap.py:6-86, map_to_crf_viterbi).
 * nread reads per call, one CTA each: read r has T_r = t_off[r+1]-t_off[r] blocks of
 * scores [T_r][S] fp32 (rows t_off[r]..), M_r = m_off[r+1]-m_off[r] positions with
 * stay_idx[m_off[r]..] and step_idx[m_off[r]-r ..] (M_r - 1 entries), traceback
 * workspace tb_ws[tb_off[r] ..] of T_r * M_r bytes.  Outputs: score[r] fp64 (position
 * scores are fp64 as in the reference), path[t_off[r]+r ..] of T_r + 1 int32 sequence
 * positions, -1 in the clipped start / end stretches (localpen).  dp_ws: 2 * sum M
 * doubles, used (and required) only when max_m positions do not fit in shared memory
 * (about 12.7k); may be NULL otherwise.  S <= 255. */
int ty_flipflop_remap(const float *scores, const int64_t *t_off, const int32_t *step_idx,
                      const int32_t *stay_idx, const int64_t *m_off, const int64_t *tb_off,
                      int nread, int S, int max_m, double localpen, double *score,
                      int32_t *path, uint8_t *tb_ws, double *dp_ws, void *stream);

/* ---------------------------------------------------------------------------
 * Dense bf16 contraction on the 5th-generation tensor cores
 * (taiyaki_b200/csrc/gemm_tc5.cu: TMA + tcgen05.mma, fp32 accumulators in tensor
 * memory).  Replaces the library GEMMs the reference reaches through nn.LSTM /
 * nn.GRU (taiyaki/layers.py:515,633: input projections and their input / weight
 * gradients inside cuDNN), nn.Conv1d (layers.py:795) and the score projection
 * nn.Linear + scale * tanh of GlobalNormFlipFlop (layers.py:1402-1411).
 *
 *     C[M x N] (fp32, row pitch ldc) = A[M x K] * B[N x K]^T
 *
 * a_mn / b_mn = 0: the operand is stored K-major ([M][lda] resp. [N][ldb] bf16);
 *             = 1: MN-major ([K][lda] resp. [K][ldb] bf16, i.e. the transpose in
 *                  place: weight gradients contract over the slow index of both).
 * Operands 16-byte aligned, lda / ldb multiples of 8.  epi:
 *   0  C = acc                          (ldc multiple of 4)
 *   1  C = scale * tanh(acc + bias[n])  (bias may be NULL)
 *   2  C[row'] += acc by red.global.add, row' = (row % map_g) * map_h + row / map_g when
 *      map_g > 0 (unit-major -> gate-major rows of the recurrent weights), else row;
 *      k_splits CTAs share an output tile
 *   3  C += acc by TMA reduce-add; k_splits as for 2
 * Asynchronous on `stream`; no workspace. */
int ty_gemm_bf16(const void *A, int lda, int a_mn, const void *B, int ldb, int b_mn, int M,
                 int N, int K, float *C, int ldc, int epi, const float *bias, float scale,
                 int k_splits, int map_g, int map_h, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TAIYAKI_B200_H */
