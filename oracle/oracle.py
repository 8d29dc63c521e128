"""oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by taiyaki_b200/).

numpy/ctypes restatement of the host side of the reference's loss operators
and a loader for the two CPU libraries:

  liboracle.so          this repo's scalar C restatement (oracle_crf.c),
                        fp32 (`f32`) and fp64 (`f64`) variants
  _ref/libctc_ref.so    the reference's own C compiled from /root/reference
                        (oracle/Makefile); present in this container and
                        shipped prebuilt to the GPU box

Each function cites the reference lines it follows.  Allowed importers:
tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke().
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SZ = ctypes.c_size_t
_FP = ctypes.POINTER(ctypes.c_float)
_ZP = ctypes.POINTER(ctypes.c_size_t)
_IP = ctypes.POINTER(ctypes.c_int32)


def build(quiet=True):
    """Compile liboracle.so (+ _ref when /root/reference is present)."""
    subprocess.run(['make', '-C', _HERE, 'all'], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)


_liboracle = None
_libref = None


def liboracle():
    global _liboracle
    if _liboracle is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            build()
        _liboracle = ctypes.CDLL(path)
    return _liboracle


def libref():
    """The reference's own C library, or None if it was never built."""
    global _libref
    if _libref is None:
        _libref = _load(os.path.join(_HERE, '_ref', 'libctc_ref.so'))
    return _libref


def have_ref():
    return libref() is not None


LARGE_VAL = 1e30          # taiyaki/constants.py:8
_libcupy = None


def libcupy_ref():
    """oracle/_ref/libcupy_ref.so: the reference's CuPy RawKernels compiled with nvcc
    (oracle/build_cupy_ref.py), or None when it was not built."""
    global _libcupy
    if _libcupy is None:
        _libcupy = _load(os.path.join(_HERE, '_ref', 'libcupy_ref.so')) or False
    return _libcupy or None


def cupy_ref_logz(scores, want_trans=True):
    """The reference's GPU path for the partition function on a CUDA tensor [T, N, S]
    (cupy_extensions/flipflop.py:88-126 flipflop_fwd, :211-246 flipflop_bwd, :299-336
    flipflop_make_trans, :338-355 LogZ): returns (logZ [N], d logZ / d scores [T, N, S]).
    The tensor set-up around the three launches is the reference wrappers'."""
    import torch
    lib = libcupy_ref()
    assert lib is not None, 'oracle/_ref/libcupy_ref.so was not built'
    c_void_p, c_ll = ctypes.c_void_p, ctypes.c_longlong
    T, N, S = scores.shape
    nbase = nbase_flipflop(S)
    scores = scores.contiguous()
    stream = c_void_p(torch.cuda.current_stream(scores.device).cuda_stream)
    fwd = torch.zeros((T + 1, N, 2 * nbase), dtype=scores.dtype, device=scores.device)
    fwd[0, :, nbase:] = -LARGE_VAL
    fwd_fact = torch.zeros((T + 1, N, 1), dtype=scores.dtype, device=scores.device)
    rc = lib.ref_cupy_flipflop_fwd(c_void_p(scores.data_ptr()), c_void_p(fwd.data_ptr()),
                                   c_void_p(fwd_fact.data_ptr()), c_ll(T), c_ll(N), c_ll(nbase), stream)
    assert rc == 0, rc
    logz = fwd_fact.sum(0)[:, 0]
    if not want_trans:
        return logz, None
    bwd = torch.zeros((T + 1, N, 2 * nbase), dtype=scores.dtype, device=scores.device)
    bwd_fact = torch.zeros((T + 1, N, 1), dtype=scores.dtype, device=scores.device)
    rc = lib.ref_cupy_flipflop_bwd(c_void_p(scores.data_ptr()), c_void_p(bwd.data_ptr()),
                                   c_void_p(bwd_fact.data_ptr()), c_ll(T), c_ll(N), c_ll(nbase), stream)
    assert rc == 0, rc
    trans = torch.zeros_like(scores)
    rc = lib.ref_cupy_flipflop_make_trans(c_void_p(scores.data_ptr()), c_void_p(fwd.data_ptr()),
                                          c_void_p(bwd.data_ptr()), c_void_p(trans.data_ptr()),
                                          c_ll(T), c_ll(N), c_ll(nbase), stream)
    assert rc == 0, rc
    return logz, trans.softmax(2)


# --------------------------------------------------------------------------
# flip-flop coding (taiyaki/flipflopfings.py)
# --------------------------------------------------------------------------
def nstate_flipflop(nbase):
    """flipflopfings.py:146-168"""
    return 2 * nbase * (nbase + 1)


def nbase_flipflop(nstate):
    """flipflopfings.py:171-184"""
    nbase_f = np.sqrt(0.25 + 0.5 * np.float32(nstate)) - 0.5
    assert np.mod(nbase_f, 1) == 0, 'Number of states not valid for flip-flop model'
    return int(np.round(nbase_f))


def move_indices(labels, nbase=4):
    """flipflopfings.py:6-17: from + min(to, nbase) * 2nbase"""
    labels = np.asarray(labels)
    return labels[:-1] + np.minimum(labels[1:], nbase) * (2 * nbase)


def stay_indices(labels, nbase=4):
    """flipflopfings.py:20-31"""
    labels = np.asarray(labels)
    return labels + np.minimum(labels, nbase) * (2 * nbase)


def flopmask(labels):
    """flipflopfings.py:34-53: True at even positions of a homopolymer run."""
    labels = np.asarray(labels)
    move = np.ediff1d(labels, to_begin=1) != 0
    cumulative = (1 - move).cumsum()
    offsets = np.maximum.accumulate(move * cumulative)
    return (cumulative - offsets) % 2 == 1


def flipflop_code(labels, alphabet_length=4):
    """flipflopfings.py:56-78"""
    x = np.array(labels).copy()
    x[flopmask(x)] += alphabet_length
    return x


def build_indices(seqs, seqlen, nbase):
    """ctc.pyx:127-132: per-chunk move/stay indices, concatenated."""
    seqs = np.asarray(seqs).astype(np.int32)
    seqlen = np.asarray(seqlen).astype(np.int32)
    parts = np.split(seqs, np.cumsum(seqlen[:-1]))
    move = np.concatenate([move_indices(s, nbase) for s in parts]).astype(np.uintp)
    stay = np.concatenate([stay_indices(s, nbase) for s in parts]).astype(np.uintp)
    return move, stay


def build_mod_indices(seqs, seqlen, mod_cats, can_mods_offsets, mod_cat_weights, nbase):
    """ctc.pyx:287-292"""
    seqs = np.asarray(seqs).astype(np.int32)
    seqlen = np.asarray(seqlen).astype(np.int32)
    mod_cats = np.asarray(mod_cats).astype(np.int32)
    can_mods_offsets = np.asarray(can_mods_offsets)
    starts = np.cumsum(seqlen[:-1])
    bs, bm = np.split(seqs, starts), np.split(mod_cats, starts)
    mod_offset = (nbase + 1) * nbase * 2
    mod_seq = np.concatenate([
        can_mods_offsets[np.mod(s[1:], nbase)] + m[1:]
        for s, m in zip(bs, bm)]).astype(int)
    modmoveidxs = (mod_offset + mod_seq).astype(np.uintp)
    modmovefacts = np.asarray(mod_cat_weights)[mod_seq].astype(np.float32)
    return modmoveidxs, modmovefacts


# --------------------------------------------------------------------------
# raw C calls (libctc.pxd:3-25 signatures)
# --------------------------------------------------------------------------
def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _pad1(a, dtype):
    """ctypes needs a valid pointer even for empty index arrays."""
    a = _c(a, dtype)
    return a if a.size else np.zeros(1, dtype=dtype)


def _names(impl):
    """impl: 'ref' (reference C), 'f32' / 'f64' (restatement)."""
    if impl == 'ref':
        lib = libref()
        assert lib is not None, 'oracle/_ref/libctc_ref.so not built'
        return lib, '{}'
    return liboracle(), 'orc_{}_' + impl


def c_crf_flipflop_grad(logprob, moveidxs, stayidxs, seqlen, impl='ref'):
    lib, fmt = _names(impl)
    lp = _c(logprob, np.float32)
    nblk, nbatch, ntrans = lp.shape
    mv, st = _pad1(moveidxs, np.uintp), _pad1(stayidxs, np.uintp)
    sl = _c(seqlen, np.int32)
    score = np.zeros(nbatch, dtype=np.float32)
    grad = np.zeros_like(lp)
    fn = getattr(lib, fmt.format('crf_flipflop_grad'))
    fn.restype = None
    fn.argtypes = [_FP, _SZ, _SZ, _SZ, _ZP, _ZP, _IP, _FP, _FP]
    fn(lp.ctypes.data_as(_FP), ntrans, nblk, nbatch, mv.ctypes.data_as(_ZP),
       st.ctypes.data_as(_ZP), sl.ctypes.data_as(_IP),
       score.ctypes.data_as(_FP), grad.ctypes.data_as(_FP))
    return score, grad


def c_crf_flipflop_cost(logprob, moveidxs, stayidxs, seqlen, impl='ref'):
    lib, fmt = _names(impl)
    lp = _c(logprob, np.float32)
    nblk, nbatch, ntrans = lp.shape
    mv, st = _pad1(moveidxs, np.uintp), _pad1(stayidxs, np.uintp)
    sl = _c(seqlen, np.int32)
    score = np.zeros(nbatch, dtype=np.float32)
    fn = getattr(lib, fmt.format('crf_flipflop_cost'))
    fn.restype = None
    fn.argtypes = [_FP, _SZ, _SZ, _SZ, _ZP, _ZP, _IP, _FP]
    fn(lp.ctypes.data_as(_FP), ntrans, nblk, nbatch, mv.ctypes.data_as(_ZP),
       st.ctypes.data_as(_ZP), sl.ctypes.data_as(_IP),
       score.ctypes.data_as(_FP))
    return score


def c_cat_mod_flipflop_grad(logprob, moveidxs, stayidxs, modmoveidxs,
                            modmovefacts, seqlen, impl='ref'):
    lib, fmt = _names(impl)
    lp = _c(logprob, np.float32)
    nblk, nbatch, ntrans = lp.shape
    mv, st = _pad1(moveidxs, np.uintp), _pad1(stayidxs, np.uintp)
    mm, mf = _pad1(modmoveidxs, np.uintp), _pad1(modmovefacts, np.float32)
    sl = _c(seqlen, np.int32)
    score = np.zeros(nbatch, dtype=np.float32)
    grad = np.zeros_like(lp)
    fn = getattr(lib, fmt.format('cat_mod_flipflop_grad'))
    fn.restype = None
    fn.argtypes = [_FP, _SZ, _SZ, _SZ, _ZP, _ZP, _ZP, _FP, _IP, _FP, _FP]
    fn(lp.ctypes.data_as(_FP), ntrans, nblk, nbatch, mv.ctypes.data_as(_ZP),
       st.ctypes.data_as(_ZP), mm.ctypes.data_as(_ZP), mf.ctypes.data_as(_FP),
       sl.ctypes.data_as(_IP), score.ctypes.data_as(_FP),
       grad.ctypes.data_as(_FP))
    return score, grad


def c_cat_mod_flipflop_cost(logprob, moveidxs, stayidxs, modmoveidxs,
                            modmovefacts, seqlen, impl='ref'):
    lib, fmt = _names(impl)
    lp = _c(logprob, np.float32)
    nblk, nbatch, ntrans = lp.shape
    mv, st = _pad1(moveidxs, np.uintp), _pad1(stayidxs, np.uintp)
    mm, mf = _pad1(modmoveidxs, np.uintp), _pad1(modmovefacts, np.float32)
    sl = _c(seqlen, np.int32)
    score = np.zeros(nbatch, dtype=np.float32)
    fn = getattr(lib, fmt.format('cat_mod_flipflop_cost'))
    fn.restype = None
    fn.argtypes = [_FP, _SZ, _SZ, _SZ, _ZP, _ZP, _ZP, _FP, _IP, _FP]
    fn(lp.ctypes.data_as(_FP), ntrans, nblk, nbatch, mv.ctypes.data_as(_ZP),
       st.ctypes.data_as(_ZP), mm.ctypes.data_as(_ZP), mf.ctypes.data_as(_FP),
       sl.ctypes.data_as(_IP), score.ctypes.data_as(_FP))
    return score


def c_scores_fb(logprob, moveidxs, stayidxs, seqlen, modmoveidxs=None,
                modmovefacts=None, impl='f32'):
    """Forward and backward scores separately -> [nbatch, 2] (restatement only)."""
    lib, fmt = _names(impl)
    lp = _c(logprob, np.float32)
    nblk, nbatch, ntrans = lp.shape
    mv, st = _pad1(moveidxs, np.uintp), _pad1(stayidxs, np.uintp)
    sl = _c(seqlen, np.int32)
    out = np.zeros((nbatch, 2), dtype=np.float32)
    fn = getattr(lib, fmt.format('crf_flipflop_scores_fb'))
    fn.restype = None
    fn.argtypes = [_FP, _SZ, _SZ, _SZ, _ZP, _ZP, _ZP, _FP, _IP, _FP]
    if modmoveidxs is None:
        mmp, mfp = None, None
    else:
        mm, mf = _pad1(modmoveidxs, np.uintp), _pad1(modmovefacts, np.float32)
        mmp, mfp = mm.ctypes.data_as(_ZP), mf.ctypes.data_as(_FP)
    fn(lp.ctypes.data_as(_FP), ntrans, nblk, nbatch, mv.ctypes.data_as(_ZP),
       st.ctypes.data_as(_ZP), mmp, mfp, sl.ctypes.data_as(_IP),
       out.ctypes.data_as(_FP))
    return out


def c_flipflop_logz(scores, want_grad=True, flop_init=-50000.0, impl='f32'):
    """logZ [N] (+ d logZ / d scores [T,N,S]) -- restatement of
    layers.py:1253-1299 and cupy_extensions/flipflop.py:128-354."""
    assert impl in ('f32', 'f64')
    lib = liboracle()
    w = _c(scores, np.float32)
    nblk, nbatch, S = w.shape
    nbase = nbase_flipflop(S)
    logz = np.zeros(nbatch, dtype=np.float32)
    grad = np.zeros_like(w) if want_grad else None
    fn = getattr(lib, 'orc_flipflop_logz_' + impl)
    fn.restype = None
    fn.argtypes = [_FP, _SZ, _SZ, _SZ, ctypes.c_float, _FP, _FP]
    fn(w.ctypes.data_as(_FP), nblk, nbatch, nbase, flop_init,
       logz.ctypes.data_as(_FP),
       grad.ctypes.data_as(_FP) if want_grad else None)
    return (logz, grad) if want_grad else logz


# --------------------------------------------------------------------------
# operator semantics (ctc.pyx:116-153 and :258-312)
# --------------------------------------------------------------------------
def crf_flipflop_loss(logprob, seqs, seqlen, sharpfact=1.0, want_grad=True,
                      impl='ref'):
    """cost[N] (= -score/nblk/sharp) and d cost / d logprob (= -G/nblk)."""
    lp = (np.float32(sharpfact) * np.asarray(logprob, dtype=np.float32))
    nblk, nbatch, ntrans = lp.shape
    nbase = nbase_flipflop(ntrans)
    move, stay = build_indices(seqs, seqlen, nbase)
    assert np.all(move < ntrans) and np.all(stay < ntrans)
    if want_grad:
        score, grad = c_crf_flipflop_grad(lp, move, stay, seqlen, impl)
        return (-score / nblk) / np.float32(sharpfact), -grad / nblk
    score = c_crf_flipflop_cost(lp, move, stay, seqlen, impl)
    return (-score / nblk) / np.float32(sharpfact)


def cat_mod_flipflop_loss(logprob, seqs, seqlen, mod_cats, can_mods_offsets,
                          mod_cat_weights, sharpfact=1.0, want_grad=True,
                          impl='ref'):
    """ctc.pyx:258-312.  Sharpening multiplies the canonical columns only;
    the returned gradient is w.r.t. the *sharpened* tensor exactly as the
    reference saves it (ctc.pyx:299), i.e. -G/nblk."""
    logprob = np.asarray(logprob, dtype=np.float32)
    nblk, nbatch, ntrans = logprob.shape
    n_can_trans = ntrans - int(can_mods_offsets[-1])
    nbase = nbase_flipflop(n_can_trans)
    trans_sharp = np.ones(ntrans, dtype=np.float32)
    trans_sharp[:n_can_trans] = sharpfact
    lp = np.ascontiguousarray(logprob * trans_sharp)
    move, stay = build_indices(seqs, seqlen, nbase)
    mm, mf = build_mod_indices(seqs, seqlen, mod_cats, can_mods_offsets,
                               mod_cat_weights, nbase)
    if want_grad:
        score, grad = c_cat_mod_flipflop_grad(lp, move, stay, mm, mf, seqlen, impl)
        return (-score / nblk) / np.float32(sharpfact), -grad / nblk
    score = c_cat_mod_flipflop_cost(lp, move, stay, mm, mf, seqlen, impl)
    return (-score / nblk) / np.float32(sharpfact)


def flipflop_logpartition(scores, want_grad=False, impl='f32'):
    """layers.py:1875-1890 -> logZ [N]"""
    return c_flipflop_logz(scores, want_grad=want_grad, impl=impl)


# --------------------------------------------------------------------------
# synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------
def flipflop_viterbi(scores):
    """Best flip-flop path, numpy restatement of taiyaki/decode.py:79-115
    (_flipflop_viterbi): fwd [T+1,N,2nb] (flip states 0, flop states -1e30), traceback
    [T,N,2nb] (best predecessor, first maximum), path [T+1,N]."""
    scores = np.asarray(scores, dtype=np.float32)
    T, N, S = scores.shape
    nb = nbase_flipflop(S)
    fwd = np.zeros((T + 1, N, 2 * nb), dtype=np.float32)
    fwd[0, :, nb:] = -1e30                       # constants.py:8 LARGE_VAL
    tb = np.zeros((T, N, 2 * nb), dtype=np.int64)
    for t in range(T):
        to_flip = scores[t, :, :S - 2 * nb].reshape(N, nb, 2 * nb)
        cand = fwd[t][:, None, :] + to_flip      # decode.py:104-106
        fwd[t + 1, :, :nb] = cand.max(2)
        tb[t, :, :nb] = cand.argmax(2)
        flop = (fwd[t] + scores[t, :, -2 * nb:]).reshape(N, 2, nb)     # decode.py:107-110
        fwd[t + 1, :, nb:] = flop.max(1)
        tb[t, :, nb:] = nb * flop.argmax(1) + np.arange(nb)
    path = np.zeros((T + 1, N), dtype=np.int64)
    path[T] = fwd[T].argmax(1)
    for t in range(T - 1, -1, -1):               # decode.py:114-116
        path[t] = tb[t, np.arange(N), path[t + 1]]
    return fwd, tb, path


def map_to_crf_viterbi(scores, step_index, stay_index, localpen=1e30):
    """Best alignment of a label sequence to transition scores, numpy restatement of
    taiyaki/flipflop_remap.py:6-86 with an explicit [T+1, M] decision table instead of
    packed bits.  Position scores are float64 (the reference's np.full default dtype),
    `start` / `end` are the clipping states that cost `localpen` per skipped block.
    Returns (score, path[T+1]) with -1 where the alignment sits in start / end."""
    # float64 throughout: what the reference computes under its pinned numpy 1.18, where the
    # scalar start / end updates promote the fp32 score like the vector updates do
    scores = np.asarray(scores, dtype=np.float64)
    step_index = np.asarray(step_index, dtype=np.int64)
    stay_index = np.asarray(stay_index, dtype=np.int64)
    T, M = len(scores), len(stay_index)
    assert len(step_index) == M - 1
    big = 1e30                                             # constants.py:8
    vec = np.full(M, -big)
    vec[0] = 0.0                                           # :31-33
    start, end, end_at = 0.0, -big, 0                      # :35-37
    moved = np.zeros((T + 1, M), dtype=np.uint8)           # :39
    for t in range(T):
        w_stay = scores[t, stay_index]                     # :44-45
        w_step = scores[t, step_index]
        stay = vec + w_stay                                # :50
        step = vec[:-1] + w_step                           # :53
        leave_start = start - localpen                     # :56
        start = start + max(w_stay[0], -localpen)          # :57
        remain = end + max(w_stay[-1], -localpen)          # :67
        into_end = vec[-1] - localpen                      # :68
        nxt = stay.copy()                                  # :60-62
        nxt[1:] = np.maximum(nxt[1:], step)
        nxt[0] = max(nxt[0], start)
        moved[t + 1, 1:] = stay[1:] < step                 # :63-64
        moved[t + 1, 0] = 1 if leave_start > stay[0] else 0
        if into_end > remain:                              # :69-71
            end_at = t
        end = max(remain, into_end)
        vec = nxt
    path = np.full(T + 1, -1, dtype=int)                   # :73
    t, m = (T, M - 1) if vec[-1] > end else (end_at, M - 1)    # :74-79
    while t >= 0 and m >= 0:                               # :81-85
        path[t] = m
        m -= int(moved[t, m])
        t -= 1
    return max(vec[-1], end), path


def remap_indices(bases, nbase=4):
    """(step_index, stay_index) of an integer base sequence (flipflop_remap.py:132-140)."""
    bases = np.asarray(bases, dtype=np.int64)
    move = np.ediff1d(bases, to_begin=1) != 0
    run = (1 - move).cumsum()
    flops = (run - np.maximum.accumulate(move * run)) % 2 == 1     # flipflopfings.py:34-53
    stay = np.where(flops, bases + (2 * nbase + 1) * nbase, bases + 2 * nbase * bases)
    frm = (bases + flops * nbase)[:-1]
    to = np.maximum(bases, nbase * flops)[1:]
    return frm + 2 * nbase * to, stay


def synth_scores(nblk, nbatch, ntrans=40, seed=0, can_nmods=None):
    """5*tanh(N(0,1)) transition scores; for cat-mod (ntrans > 40) the extra
    columns are per-canonical-base log-softmax groups (layers.py:1611-1640)."""
    rng = np.random.RandomState(seed)
    ncan = 40 if ntrans > 40 else ntrans
    x = (5.0 * np.tanh(rng.standard_normal((nblk, nbatch, ncan)))).astype(np.float32)
    if ntrans == ncan:
        return x
    can_nmods = [0, 1, 0, 0] if can_nmods is None else can_nmods
    cols = []
    for nm in can_nmods:
        z = rng.standard_normal((nblk, nbatch, nm + 1))
        z = z - np.log(np.exp(z).sum(-1, keepdims=True))
        cols.append(z)
    return np.concatenate([x] + cols, axis=2).astype(np.float32)


def synth_seqs(nblk, nbatch, stride=5, seed=1, nbase=4, samples_per_base=9.0,
               lengths=None):
    """Random base sequences, flip-flop coded, L_b = round(nblk*stride/9*u),
    u~U[0.9,1.1] (r9.4.1-like dwell).  Returns (seqs int64, seqlen int64,
    raw base labels)."""
    rng = np.random.RandomState(seed)
    if lengths is None:
        lengths = np.maximum(1, np.round(
            nblk * stride / samples_per_base * rng.uniform(0.9, 1.1, nbatch))
        ).astype(np.int64)
        lengths = np.minimum(lengths, nblk)
    lengths = np.asarray(lengths, dtype=np.int64)
    raw = [rng.randint(0, nbase, size=int(L)) for L in lengths]
    seqs = [flipflop_code(r, nbase) if len(r) else r for r in raw]
    cat = np.concatenate(seqs).astype(np.int64) if len(seqs) else np.zeros(0, np.int64)
    return cat, lengths, raw
