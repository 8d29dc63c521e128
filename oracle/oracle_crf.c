/*
 * oracle_crf.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the
 * flip-flop CRF training dynamic programs of nanoporetech/taiyaki v5.3.0.
 *
 * Nothing under oracle/ is on the product path: only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * it, and only as the checker.
 *
 * Plain scalar C (no AVX, no OpenMP): every function states which reference
 * lines it follows.  The real type is a macro so the same text compiles as
 * the fp32 restatement (REAL=float, what the reference computes in) and as an
 * fp64 "ground truth" used to rank rounding errors (REAL=double).
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against
 *   - the reference's own C compiled from /root/reference (oracle/_ref), on
 *     the embedded 7x2x40 / 7x2x45 tables of c_crf_flipflop.c:520-695 and
 *     c_cat_mod_flipflop.c:586-870 (fwd = bwd = -2.378088, -52.354622 /
 *     -195.435257) and on seeded random batches,
 *   - golden vectors generated from the reference (tests/golden/),
 *   - test/unit/test_ctc_loss.py:39-103 path probabilities.
 *
 * Signatures are the reference's (taiyaki/ctc/libctc.pxd:3-25) with an
 * `orc_` prefix, so one ctypes wrapper drives either library.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#define LARGE_VAL ((REAL)1e30)          /* c_crf_flipflop.c:11 */

static inline REAL r_exp(REAL x) { return (REAL)exp((double)x); }
static inline REAL r_log(REAL x) { return (REAL)log((double)x); }

/* vect_mathfun.h:79-102: max(x,y) + log(1 + exp(-|x-y|)) */
static inline REAL logaddexp_r(REAL x, REAL y) {
    REAL mx = x > y ? x : y;
    REAL d = x - y;
    if (d < 0) d = -d;
    return mx + r_log((REAL)1 + r_exp(-d));
}

/* One forward step, c_crf_flipflop.c:43-78 (plain) and
 * c_cat_mod_flipflop.c:37-78 (mod term on every move).  Returns the
 * normalisation factor (max over positions). */
static REAL forward_step(const REAL *lp, const REAL *prev, const size_t *mv,
                         const size_t *st, const size_t *mmv, const float *mfact,
                         size_t L, REAL *cur) {
    REAL f = -HUGE_VAL;
    for (size_t p = 0; p < L; p++) {
        REAL stay = lp[st[p]] + prev[p];
        REAL v;
        if (p == 0) {
            /* fwdtmp[0] = -inf: logaddexp(-inf, stay) = stay */
            v = stay;
        } else {
            REAL move = prev[p - 1] + lp[mv[p - 1]];
            if (mmv) move += lp[mmv[p - 1]] * (REAL)mfact[p - 1];
            v = logaddexp_r(move, stay);
        }
        cur[p] = v;
        if (v > f) f = v;
    }
    for (size_t p = 0; p < L; p++) cur[p] -= f;
    return f;
}

/* One backward step, c_crf_flipflop.c:150-182 / c_cat_mod_flipflop.c:163-196 */
static REAL backward_step(const REAL *lp, const REAL *prev, const size_t *mv,
                          const size_t *st, const size_t *mmv,
                          const float *mfact, size_t L, REAL *cur) {
    REAL f = -HUGE_VAL;
    for (size_t p = 0; p < L; p++) {
        REAL stay = lp[st[p]] + prev[p];
        REAL v;
        if (p == L - 1) {
            v = stay;
        } else {
            REAL move = prev[p + 1] + lp[mv[p]];
            if (mmv) move += lp[mmv[p]] * (REAL)mfact[p];
            v = logaddexp_r(move, stay);
        }
        cur[p] = v;
        if (v > f) f = v;
    }
    for (size_t p = 0; p < L; p++) cur[p] -= f;
    return f;
}

/* c_crf_flipflop.c:97-133: alpha_0 = [0, -1e30...]; score = sum f + alpha_T[L-1].
 * `row` is a scratch row of ntrans REALs holding the (converted) score row. */
static REAL forward_all(const float *logprob, size_t ntrans, size_t nblk,
                        size_t ldp, const size_t *mv, const size_t *st,
                        const size_t *mmv, const float *mfact, size_t L,
                        REAL *fwd, REAL *row) {
    for (size_t p = 0; p < L; p++) fwd[p] = -LARGE_VAL;
    fwd[0] = 0;
    REAL score = 0;
    for (size_t blk = 0; blk < nblk; blk++) {
        for (size_t s = 0; s < ntrans; s++) row[s] = (REAL)logprob[blk * ldp + s];
        score += forward_step(row, fwd + blk * L, mv, st, mmv, mfact, L,
                              fwd + (blk + 1) * L);
    }
    return score + fwd[nblk * L + L - 1];
}

/* c_crf_flipflop.c:201-235: beta_T = [-1e30..., 0]; score = beta_0[0] + sum g */
static REAL backward_all(const float *logprob, size_t ntrans, size_t nblk,
                         size_t ldp, const size_t *mv, const size_t *st,
                         const size_t *mmv, const float *mfact, size_t L,
                         REAL *bwd, REAL *row) {
    for (size_t p = 0; p < L; p++) bwd[nblk * L + p] = -LARGE_VAL;
    bwd[nblk * L + L - 1] = 0;
    REAL score = 0;
    for (size_t blk = nblk; blk > 0; blk--) {
        for (size_t s = 0; s < ntrans; s++)
            row[s] = (REAL)logprob[(blk - 1) * ldp + s];
        score += backward_step(row, bwd + blk * L, mv, st, mmv, mfact, L,
                               bwd + (blk - 1) * L);
    }
    return bwd[0] + score;
}

/* Posterior of one block, c_crf_flipflop.c:372-413 and
 * c_cat_mod_flipflop.c:419-468: softmax over the 2L-1 joint stay/move scores,
 * scatter-added into the ntrans bins. */
static void grad_step(const REAL *fwdcur, const REAL *bwdnext, const REAL *lp,
                      const size_t *mv, const size_t *st, const size_t *mmv,
                      const float *mfact, size_t L, float *grad, REAL *tmp,
                      size_t ntrans, REAL *acc) {
    for (size_t s = 0; s < ntrans; s++) acc[s] = 0;
    REAL mx = -HUGE_VAL;
    for (size_t p = 0; p < L; p++) {
        tmp[p] = fwdcur[p] + bwdnext[p] + lp[st[p]];
        if (tmp[p] > mx) mx = tmp[p];
    }
    for (size_t p = 0; p + 1 < L; p++) {
        REAL v = fwdcur[p] + bwdnext[p + 1] + lp[mv[p]];
        if (mmv) v += lp[mmv[p]] * (REAL)mfact[p];
        tmp[L + p] = v;
        if (v > mx) mx = v;
    }
    const size_t n = 2 * L - 1;
    REAL Z = 0;
    for (size_t i = 0; i < n; i++) {
        tmp[i] = r_exp(tmp[i] - mx);
        Z += tmp[i];
    }
    for (size_t i = 0; i < n; i++) tmp[i] /= Z;
    for (size_t p = 0; p < L; p++) acc[st[p]] += tmp[p];
    for (size_t p = 0; p + 1 < L; p++) {
        acc[mv[p]] += tmp[L + p];
        if (mmv) acc[mmv[p]] += tmp[L + p] * (REAL)mfact[p];
    }
    for (size_t s = 0; s < ntrans; s++) grad[s] = (float)acc[s];
}

/* Batch drivers: c_crf_flipflop.c:255-290 (cost), :434-516 (grad);
 * c_cat_mod_flipflop.c:286-345, :493-582.  Index packing: stay arrays are
 * offset by sum(seqlen[:b]), move (and mod) arrays by that minus b. */
static void batch_cost(const float *logprob, size_t ntrans, size_t nblk,
                       size_t nbatch, const size_t *moveidxs,
                       const size_t *stayidxs, const size_t *modmoveidxs,
                       const float *modmovefacts, const int32_t *seqlen,
                       float *score) {
    const size_t ldp = nbatch * ntrans;
    size_t seqidx = 0;
    for (size_t b = 0; b < nbatch; b++) {
        const size_t L = (size_t)seqlen[b];
        if (L == 0) {
            score[b] = 0.0f;
            continue;
        }
        REAL *fwd = malloc((nblk + 1) * L * sizeof(REAL));
        REAL *row = malloc(ntrans * sizeof(REAL));
        score[b] = (float)forward_all(
            logprob + b * ntrans, ntrans, nblk, ldp, moveidxs + seqidx - b,
            stayidxs + seqidx, modmoveidxs ? modmoveidxs + seqidx - b : NULL,
            modmovefacts ? modmovefacts + seqidx - b : NULL, L, fwd, row);
        free(row);
        free(fwd);
        seqidx += L;
    }
}

static void batch_grad(const float *logprob, size_t ntrans, size_t nblk,
                       size_t nbatch, const size_t *moveidxs,
                       const size_t *stayidxs, const size_t *modmoveidxs,
                       const float *modmovefacts, const int32_t *seqlen,
                       float *score, float *grad, float *score_fb) {
    const size_t ldp = nbatch * ntrans;
    size_t seqidx = 0;
    for (size_t b = 0; b < nbatch; b++) {
        const size_t L = (size_t)seqlen[b];
        if (L == 0) {
            /* c_crf_flipflop.c:458-464: gradient rows zeroed; score untouched
             * (caller passes a zero-filled vector, ctc.pyx:96) */
            for (size_t blk = 0; blk < nblk; blk++)
                memset(grad + b * ntrans + blk * ldp, 0, ntrans * sizeof(float));
            continue;
        }
        const size_t *mv = moveidxs + seqidx - b;
        const size_t *st = stayidxs + seqidx;
        const size_t *mmv = modmoveidxs ? modmoveidxs + seqidx - b : NULL;
        const float *mf = modmovefacts ? modmovefacts + seqidx - b : NULL;
        REAL *fwd = malloc((nblk + 1) * L * sizeof(REAL));
        REAL *bwd = malloc((nblk + 1) * L * sizeof(REAL));
        REAL *tmp = malloc(2 * L * sizeof(REAL));
        REAL *row = malloc(2 * ntrans * sizeof(REAL));
        const float *lpb = logprob + b * ntrans;
        REAL F = forward_all(lpb, ntrans, nblk, ldp, mv, st, mmv, mf, L, fwd, row);
        REAL B = backward_all(lpb, ntrans, nblk, ldp, mv, st, mmv, mf, L, bwd, row);
        score[b] = (float)((F + B) * (REAL)0.5);        /* :482-491 */
        if (score_fb) {
            score_fb[2 * b] = (float)F;
            score_fb[2 * b + 1] = (float)B;
        }
        for (size_t blk = 0; blk < nblk; blk++) {
            for (size_t s = 0; s < ntrans; s++)
                row[s] = (REAL)lpb[blk * ldp + s];
            grad_step(fwd + blk * L, bwd + (blk + 1) * L, row, mv, st, mmv, mf,
                      L, grad + b * ntrans + blk * ldp, tmp, ntrans,
                      row + ntrans);
        }
        free(row);
        free(tmp);
        free(bwd);
        free(fwd);
        seqidx += L;
    }
}

void FN(orc_crf_flipflop_cost)(const float *logprob, size_t ntrans, size_t nblk,
                               size_t nbatch, const size_t *moveidxs,
                               const size_t *stayidxs, const int32_t *seqlen,
                               float *score) {
    batch_cost(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, NULL, NULL,
               seqlen, score);
}

void FN(orc_crf_flipflop_grad)(const float *logprob, size_t ntrans, size_t nblk,
                               size_t nbatch, const size_t *moveidxs,
                               const size_t *stayidxs, const int32_t *seqlen,
                               float *score, float *grad) {
    batch_grad(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, NULL, NULL,
               seqlen, score, grad, NULL);
}

void FN(orc_cat_mod_flipflop_cost)(const float *logprob, size_t ntrans,
                                   size_t nblk, size_t nbatch,
                                   const size_t *moveidxs,
                                   const size_t *stayidxs,
                                   const size_t *modmoveidxs,
                                   const float *modmovefacts,
                                   const int32_t *seqlen, float *score) {
    batch_cost(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, modmoveidxs,
               modmovefacts, seqlen, score);
}

void FN(orc_cat_mod_flipflop_grad)(const float *logprob, size_t ntrans,
                                   size_t nblk, size_t nbatch,
                                   const size_t *moveidxs,
                                   const size_t *stayidxs,
                                   const size_t *modmoveidxs,
                                   const float *modmovefacts,
                                   const int32_t *seqlen, float *score,
                                   float *grad) {
    batch_grad(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, modmoveidxs,
               modmovefacts, seqlen, score, grad, NULL);
}

/* Forward and backward scores separately (c_crf_flipflop.c:297-365); used to
 * pin F == B on the reference's embedded tables. score_fb is [nbatch][2]. */
void FN(orc_crf_flipflop_scores_fb)(const float *logprob, size_t ntrans,
                                    size_t nblk, size_t nbatch,
                                    const size_t *moveidxs,
                                    const size_t *stayidxs,
                                    const size_t *modmoveidxs,
                                    const float *modmovefacts,
                                    const int32_t *seqlen, float *score_fb) {
    float *score = calloc(nbatch, sizeof(float));
    float *grad = calloc(nblk * nbatch * ntrans, sizeof(float));
    batch_grad(logprob, ntrans, nblk, nbatch, moveidxs, stayidxs, modmoveidxs,
               modmovefacts, seqlen, score, grad, score_fb);
    free(grad);
    free(score);
}

/* ------------------------------------------------------------------------
 * Partition function over the 2*nbase-state lattice (all paths).
 *
 * Forward follows taiyaki/layers.py:1253-1299 (log_partition_flipflop /
 * global_norm_flipflop_step): phi_0 = [0]*nbase + [flop_init]*nbase,
 * normalised; per step
 *    flip t': logsumexp_f(phi[f] + w[t'*2nbase + f])
 *    flop b : logaddexp(phi[b] + w[2nbase*nbase + b],
 *                       phi[nbase+b] + w[2nbase*nbase + nbase + b])
 * then row-normalise, logZ accumulates the factors.
 * flop_init = -50000 (layers.py:1289, LARGE_LOG_VAL) or -1e30
 * (cupy_extensions/flipflop.py:117-118) -- identical logZ in fp32.
 *
 * Backward + posterior follow cupy_extensions/flipflop.py:128-208 (psi_T =
 * -log(2 nbase), free end), :248-296 (trans = fwd[t,from] + w + bwd[t+1,to])
 * and :338-354 (gradient = softmax over the S transitions of a block).
 * ---------------------------------------------------------------------- */
static REAL lse_vec(const REAL *x, size_t n) {
    REAL mx = x[0];
    for (size_t i = 1; i < n; i++) if (x[i] > mx) mx = x[i];
    REAL z = 0;
    for (size_t i = 0; i < n; i++) z += r_exp(x[i] - mx);
    return mx + r_log(z);
}

void FN(orc_flipflop_logz)(const float *scores, size_t nblk, size_t nbatch,
                           size_t nbase, float flop_init, float *logz,
                           float *grad /* may be NULL */) {
    const size_t ns = 2 * nbase;
    const size_t S = 2 * nbase * (nbase + 1);
    const size_t ld = nbatch * S;
    REAL *fwd = malloc((nblk + 1) * ns * sizeof(REAL));
    REAL *bwd = malloc((nblk + 1) * ns * sizeof(REAL));
    REAL *tmp = malloc((S > ns ? S : ns) * sizeof(REAL));
    for (size_t b = 0; b < nbatch; b++) {
        const float *w = scores + b * S;
        REAL *phi = fwd;
        for (size_t s = 0; s < nbase; s++) phi[s] = 0;
        for (size_t s = nbase; s < ns; s++) phi[s] = (REAL)flop_init;
        REAL lz = lse_vec(phi, ns);
        for (size_t s = 0; s < ns; s++) phi[s] -= lz;
        for (size_t t = 0; t < nblk; t++) {
            const float *wt = w + t * ld;
            const REAL *prev = fwd + t * ns;
            REAL *cur = fwd + (t + 1) * ns;
            for (size_t to = 0; to < nbase; to++) {
                for (size_t f = 0; f < ns; f++)
                    tmp[f] = prev[f] + (REAL)wt[to * ns + f];
                cur[to] = lse_vec(tmp, ns);
                cur[nbase + to] = logaddexp_r(
                    prev[to] + (REAL)wt[ns * nbase + to],
                    prev[nbase + to] + (REAL)wt[ns * nbase + nbase + to]);
            }
            REAL fac = lse_vec(cur, ns);
            for (size_t s = 0; s < ns; s++) cur[s] -= fac;
            lz += fac;
        }
        logz[b] = (float)lz;
        if (!grad) continue;
        REAL *psi = bwd + nblk * ns;
        for (size_t s = 0; s < ns; s++) psi[s] = -r_log((REAL)ns);
        for (size_t t = nblk; t > 0; t--) {
            const float *wt = w + (t - 1) * ld;
            const REAL *nxt = bwd + t * ns;
            REAL *cur = bwd + (t - 1) * ns;
            for (size_t f = 0; f < ns; f++) {
                size_t n = 0;
                for (size_t to = 0; to < nbase; to++)
                    tmp[n++] = nxt[to] + (REAL)wt[to * ns + f];
                size_t toflop = f < nbase ? f + nbase : f;
                tmp[n++] = nxt[toflop] + (REAL)wt[ns * nbase + f];
                cur[f] = lse_vec(tmp, n);
            }
            REAL fac = lse_vec(cur, ns);
            for (size_t s = 0; s < ns; s++) cur[s] -= fac;
        }
        for (size_t t = 0; t < nblk; t++) {
            const float *wt = w + t * ld;
            const REAL *f_ = fwd + t * ns;
            const REAL *b_ = bwd + (t + 1) * ns;
            for (size_t f = 0; f < ns; f++) {
                for (size_t to = 0; to < nbase; to++)
                    tmp[to * ns + f] = f_[f] + (REAL)wt[to * ns + f] + b_[to];
                size_t toflop = f < nbase ? f + nbase : f;
                tmp[ns * nbase + f] =
                    f_[f] + (REAL)wt[ns * nbase + f] + b_[toflop];
            }
            REAL l = lse_vec(tmp, S);
            float *g = grad + b * S + t * ld;
            for (size_t s = 0; s < S; s++) g[s] = (float)r_exp(tmp[s] - l);
        }
    }
    free(tmp);
    free(bwd);
    free(fwd);
}
