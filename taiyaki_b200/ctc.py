"""Flip-flop CRF loss operators on the device -- drop-in for taiyaki/ctc/ctc.pyx.

Same names, positional arguments, return shapes and gradient conventions as the
reference (`crf_flipflop_loss` ctc.pyx:153, `cat_mod_flipflop_loss` ctc.pyx:312),
but the score tensor never leaves HBM: indices are built on the device
(ctc.pyx:127-132 -> ty_flipflop_indices) and the forward/backward/posterior DP
runs in the kernels of csrc/crf_flipflop.cu (c_crf_flipflop.c:434-516).

As in the reference the gradient is computed eagerly in `forward` when
`logprob.requires_grad`, and `backward` is one broadcast multiply
(ctc.pyx:136-151).
"""
import numpy as np
import torch

from . import _lib
from . import flipflopfings

#: When True every call checks costs / gradients for non-finite values and
#: raises the reference's AssertionError (ctc.pyx:62-65, :107-112).  This forces
#: a device synchronisation per call, so the training loop leaves it off and
#: checks the scalar loss it reads back anyway.
CHECK_FINITE = False

_FINITE_MSG = ("Error: all costs must be finite, got {}.\n"
               "Try restarting from a checkpoint with a lower learning rate.")


def nstate_to_nbase(nstate):
    """Number of bases for `nstate` transitions, asserting validity (ctc.pyx:13-21)."""
    nbase_f = np.sqrt(0.25 + (0.5 * nstate)) - 0.5
    assert np.mod(nbase_f, 1) == 0, (
        'Number of states not valid for flip-flop model. ' +
        'nstates: {}\tconverted nbases: {}').format(nstate, nbase_f)
    return int(nbase_f)


#: Lazy input checks.  ctc.pyx:133-134 asserts label indices on the host, which here would
#: cost a device synchronisation per call.  Instead the index kernel clamps bad labels (so
#: nothing gathers out of bounds) and ORs into a per-device flag word
#: (ty_flipflop_indices_checked); training.TrainStep copies the flag back with the loss of
#: the step, `check_pending()` does it on demand; both raise the reference's AssertionError.
_FLAGS = {}          # device index -> int32[1], ORed into by ty_flipflop_indices_checked


def _flag_tensor(device):
    t = _FLAGS.get(device.index)
    if t is None:
        t = _FLAGS[device.index] = torch.zeros(1, dtype=torch.int32, device=device)
    return t


def pending_flags(device=None):
    """Snapshot of the label-range flag as a device tensor (None if no operator ran on
    that device yet); the flag is cleared."""
    if device is None:
        if not _FLAGS:
            return None
        device = torch.device('cuda', next(iter(_FLAGS)))
    t = _FLAGS.get(device.index)
    if t is None:
        return None
    snap = t.clone()
    t.zero_()
    return snap


def raise_if_flagged(flag_value):
    assert not flag_value, 'Error: label indices out of range for the flip-flop model (ctc.pyx:133-134)'


def check_pending():
    for idx in list(_FLAGS):
        flag = pending_flags(torch.device('cuda', idx))
        raise_if_flagged(bool(flag.item()))


def _as_device_i64(x, device):
    if not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x))
    return x.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


def hint_lengths(seqlen, max_len, total):
    """Tell the ops the (max, sum) of a DEVICE-resident seqlen tensor so they
    need not synchronise to read it back.  The hint rides on the tensor object
    itself, so it is released with the batch (a module-level table keyed by id()
    kept every batch's tensor alive for the whole run)."""
    seqlen._ty_len_hint = (int(max_len), int(total))
    return seqlen


def _max_len(seqlen):
    """Longest sequence; free when seqlen lives on the host (the reference's
    batching hands over CPU tensors, train_flipflop.py:133-135)."""
    hint = getattr(seqlen, '_ty_len_hint', None)
    if hint is not None:
        return hint
    if not torch.is_tensor(seqlen):
        seqlen = torch.as_tensor(np.asarray(seqlen))
    return int(seqlen.max()) if seqlen.numel() else 0, int(seqlen.sum())


_SMALL_ARRAYS = {}


def _small_device_array(values, dtype, device):
    """Device copy of a small host array (mod offsets / weights), cached by value: the same few
    arrays come back every step, and a host -> device copy per call is neither free nor allowed
    while a CUDA graph is being captured."""
    arr = np.ascontiguousarray(np.asarray(values, dtype=dtype))
    key = (arr.tobytes(), arr.dtype.str, str(device))
    t = _SMALL_ARRAYS.get(key)
    if t is None:
        if len(_SMALL_ARRAYS) > 256:        # mod factors ramp during training: bounded cache
            _SMALL_ARRAYS.clear()
        t = _SMALL_ARRAYS[key] = torch.from_numpy(arr.copy()).to(device)
    return t


def build_indices(seqs, seqlen, nbase, device, mod_cats=None, can_mods_offsets=None,
                  mod_cat_weights=None):
    """Device-side move/stay (and mod) transition indices in the reference packing."""
    lib = _lib.lib()
    max_len, total = _max_len(seqlen)
    nbatch = int(len(seqlen))
    seqs_d = _as_device_i64(seqs, device)
    seqlen_d = _as_device_i64(seqlen, device)
    assert seqs_d.numel() == total, 'sum(seqlen) != len(seqs)'
    n = max(total, 1)
    move = torch.zeros(n, dtype=torch.int32, device=device)
    stay = torch.zeros(n, dtype=torch.int32, device=device)
    seqlen32 = torch.empty(nbatch, dtype=torch.int32, device=device)
    modmove = modfact = mod_d = off_d = w_d = None
    if mod_cats is not None:
        mod_d = _as_device_i64(mod_cats, device)
        off_d = _small_device_array(can_mods_offsets, np.int32, device)
        w_d = _small_device_array(mod_cat_weights, np.float32, device)
        modmove = torch.zeros(n, dtype=torch.int32, device=device)
        modfact = torch.zeros(n, dtype=torch.float32, device=device)
    rc = lib.ty_flipflop_indices_checked(
        _lib.ptr(seqs_d), _lib.ptr(seqlen_d), nbatch, total, nbase, _lib.ptr(mod_d),
        _lib.ptr(off_d), _lib.ptr(w_d), _lib.ptr(move), _lib.ptr(stay), _lib.ptr(seqlen32),
        _lib.ptr(modmove), _lib.ptr(modfact), _lib.ptr(_flag_tensor(device)),
        _lib.stream_ptr(device))
    _lib.check(rc, 'ty_flipflop_indices')
    _lib.count_launches(1)
    return move, stay, seqlen32, modmove, modfact, max_len


def _run_crf(logprob, move, stay, modmove, modfact, seqlen32, max_len, sharpfact, nsharp,
             want_grad):
    lib = _lib.lib()
    device = logprob.device
    lp = logprob.detach()
    if lp.dtype != torch.float32 or not lp.is_contiguous():
        lp = lp.float().contiguous()
    nblk, nbatch, ntrans = lp.shape
    cost = torch.empty(nbatch, dtype=torch.float32, device=device)
    grads = torch.empty_like(lp) if want_grad else None
    ws_bytes = lib.ty_crf_flipflop_workspace_bytes(ntrans, nblk, nbatch, max_len, int(want_grad))
    ws = _lib.workspace(ws_bytes, device)
    # cost = -score / nblk / sharp, grad = -G / nblk   (ctc.pyx:66,113,145)
    with _lib.timed('crf_fwd_bwd' if want_grad else 'crf_fwd', device):
        rc = lib.ty_crf_flipflop(
            _lib.ptr(lp), ntrans, nblk, nbatch, _lib.ptr(move), _lib.ptr(stay),
            _lib.ptr(modmove), _lib.ptr(modfact), _lib.ptr(seqlen32), max_len, float(sharpfact),
            nsharp, -1.0 / (nblk * float(sharpfact)), _lib.ptr(cost), -1.0 / nblk,
            _lib.ptr(grads), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(device))
    _lib.check(rc, 'ty_crf_flipflop')
    _lib.count_launches(2 if lib.ty_crf_last_path() == 1 else 1)   # chain + posterior kernels, or one fused kernel
    if CHECK_FINITE:
        assert bool(torch.isfinite(cost).all()), _FINITE_MSG.format(cost.cpu().numpy())
        if grads is not None:
            assert bool(torch.isfinite(grads).all()), (
                "Error: Gradients not finite.\n"
                "Try restarting from a checkpoint with a lower learning rate.")
    return cost, grads


class FlipFlopCRF(torch.autograd.Function):
    """Device implementation of ctc.pyx:116-151."""

    @staticmethod
    def forward(ctx, logprob, seqs, seqlen, sharpfact):
        _lib.require_cuda(logprob, 'logprob')
        ntrans = logprob.shape[2]
        nbase = flipflopfings.nbase_flipflop(ntrans)
        move, stay, seqlen32, _, _, max_len = build_indices(seqs, seqlen, nbase, logprob.device)
        cost, grads = _run_crf(logprob, move, stay, None, None, seqlen32, max_len, sharpfact,
                               ntrans, logprob.requires_grad)
        if grads is not None:
            ctx.save_for_backward(grads)
        return cost

    @staticmethod
    def backward(ctx, output_grads):
        grads, = ctx.saved_tensors
        return grads * output_grads.unsqueeze(1), None, None, None


crf_flipflop_loss = FlipFlopCRF.apply


class CatModFlipFlop(torch.autograd.Function):
    """Device implementation of ctc.pyx:258-310 (categorical modified bases)."""

    @staticmethod
    def forward(ctx, logprob, seqs, seqlen, mod_cats, can_mods_offsets, mod_cat_weights,
                sharpfact):
        _lib.require_cuda(logprob, 'logprob')
        ntrans = logprob.shape[2]
        n_can_trans = ntrans - int(can_mods_offsets[-1])
        nbase = flipflopfings.nbase_flipflop(n_can_trans)
        move, stay, seqlen32, modmove, modfact, max_len = build_indices(
            seqs, seqlen, nbase, logprob.device, mod_cats, can_mods_offsets, mod_cat_weights)
        # sharpening multiplies the canonical transition columns only (ctc.pyx:265-269)
        cost, grads = _run_crf(logprob, move, stay, modmove, modfact, seqlen32, max_len,
                               sharpfact, n_can_trans, logprob.requires_grad)
        if grads is not None:
            ctx.save_for_backward(grads)
        return cost

    @staticmethod
    def backward(ctx, output_grads):
        grads, = ctx.saved_tensors
        return (grads * output_grads.unsqueeze(1), None, None, None, None, None, None)


cat_mod_flipflop_loss = CatModFlipFlop.apply


def crf_flipflop_cost_grad(logprob, seqs, seqlen, sharpfact=1.0, want_grad=True):
    """Functional form returning (cost, dcost/dlogprob) without autograd."""
    ntrans = logprob.shape[2]
    nbase = flipflopfings.nbase_flipflop(ntrans)
    move, stay, seqlen32, _, _, max_len = build_indices(seqs, seqlen, nbase, logprob.device)
    return _run_crf(logprob, move, stay, None, None, seqlen32, max_len, sharpfact, ntrans,
                    want_grad)


class FlipFlopTrainLoss(torch.autograd.Function):
    """cost + logZ / nblk of bin/train_flipflop.py:163-182 as one operator
    (ty_flipflop_train_loss): the label-constrained and the partition-function
    chains run concurrently and the combined gradient is written once.
    Numerically the sum of `crf_flipflop_loss` (or `cat_mod_flipflop_loss`) and
    `layers.flipflop_logpartition(outputs[:, :, :40]) / nblk`."""

    @staticmethod
    def forward(ctx, logprob, seqs, seqlen, sharpfact, mod_cats, can_mods_offsets,
                mod_cat_weights):
        _lib.require_cuda(logprob, 'logprob')
        lib = _lib.lib()
        device = logprob.device
        lp = logprob.detach()
        if lp.dtype != torch.float32 or not lp.is_contiguous():
            lp = lp.float().contiguous()
        nblk, nbatch, ntrans = lp.shape
        ncan = ntrans if mod_cats is None else ntrans - int(can_mods_offsets[-1])
        nbase = flipflopfings.nbase_flipflop(ncan)
        move, stay, seqlen32, modmove, modfact, max_len = build_indices(
            seqs, seqlen, nbase, device, mod_cats, can_mods_offsets, mod_cat_weights)
        want_grad = logprob.requires_grad
        cost = torch.empty(nbatch, dtype=torch.float32, device=device)
        logz = torch.empty(nbatch, dtype=torch.float32, device=device)
        grads = torch.empty_like(lp) if want_grad else None
        ws_bytes = lib.ty_flipflop_train_loss_workspace_bytes(ntrans, nblk, nbatch, max_len,
                                                              int(want_grad))
        ws = _lib.workspace(ws_bytes, device)
        with _lib.timed('loss_fwd_bwd' if want_grad else 'loss_fwd', device):
            rc = lib.ty_flipflop_train_loss(
                _lib.ptr(lp), ntrans, nblk, nbatch, _lib.ptr(move), _lib.ptr(stay),
                _lib.ptr(modmove), _lib.ptr(modfact), _lib.ptr(seqlen32), max_len,
                float(sharpfact), ncan, _lib.ptr(cost), _lib.ptr(logz), _lib.ptr(grads),
                _lib.ptr(ws), ws.numel(), _lib.stream_ptr(device))
        _lib.check(rc, 'ty_flipflop_train_loss')
        # logZ chains + logZ posterior + CRF chains (+ CRF posterior kernel unless it is fused into the chains)
        _lib.count_launches((3 + (lib.ty_crf_last_path() == 1)) if want_grad else 2)
        if want_grad:
            ctx.save_for_backward(grads)
        return cost + logz

    @staticmethod
    def backward(ctx, output_grads):
        grads, = ctx.saved_tensors
        return (grads * output_grads.unsqueeze(1), None, None, None, None, None, None)


def flipflop_train_loss(logprob, seqs, seqlen, sharpfact, mod_cats=None, can_mods_offsets=None,
                        mod_cat_weights=None):
    """Per-chunk training loss vector [N] = CRF cost + logZ / nblk."""
    return FlipFlopTrainLoss.apply(logprob, seqs, seqlen, sharpfact, mod_cats, can_mods_offsets,
                                   mod_cat_weights)
