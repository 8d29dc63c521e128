"""Inference lattice operators of taiyaki/decode.py on the device:
`flipflop_viterbi` (:15-41, :79-115) and `flipflop_make_trans` (:44-76), same
signatures and return values; `_never_use_cupy` is accepted for compatibility
(there is one device implementation)."""
import torch

from . import _lib, flipflopfings, layers


@torch.no_grad()
def flipflop_viterbi(scores, _never_use_cupy=False):
    """Highest scoring flip-flop paths for [T, N, S] scores.  Returns (fwd scores
    [T+1, N, 2 nbase] fp32, traceback [T, N, 2 nbase] int64, path [T+1, N] int64)."""
    _lib.require_cuda(scores, 'scores')
    lib = _lib.lib()
    T, N, S = scores.shape
    nbase = flipflopfings.nbase_flipflop(S)
    x = scores.detach().float().contiguous()
    dev = x.device
    fwd = torch.empty(T + 1, N, 2 * nbase, dtype=torch.float32, device=dev)
    traceback = torch.empty(T, N, 2 * nbase, dtype=torch.int64, device=dev)
    path = torch.empty(T + 1, N, dtype=torch.int64, device=dev)
    rc = lib.ty_flipflop_viterbi(_lib.ptr(x), T, N, nbase, _lib.ptr(fwd), _lib.ptr(traceback),
                                 _lib.ptr(path), _lib.stream_ptr(dev))
    _lib.check(rc, 'ty_flipflop_viterbi')
    _lib.count_launches(1)
    return fwd, traceback, path


@torch.no_grad()
def flipflop_make_trans(scores, _never_use_cupy=False):
    """Posterior transition probabilities (not logs) of [T, N, S] scores: the
    derivative of the log-partition function with respect to the scores, which
    is what the partition-function posterior kernel writes.  Called directly (no
    autograd graph), so it behaves the same inside `torch.no_grad()` blocks."""
    _lib.require_cuda(scores, 'scores')
    lib = _lib.lib()
    T, N, S = scores.shape
    nbase = flipflopfings.nbase_flipflop(S)
    x = scores.detach().float().contiguous()
    dev = x.device
    logz = torch.empty(N, dtype=torch.float32, device=dev)
    trans = torch.empty(T, N, S, dtype=torch.float32, device=dev)
    ws = _lib.workspace(lib.ty_flipflop_logz_workspace_bytes(nbase, T, N), dev)
    rc = lib.ty_flipflop_logz(_lib.ptr(x), S, T, N, nbase, 1.0, _lib.ptr(logz), 1.0,
                              _lib.ptr(trans), S, 0, _lib.ptr(ws), ws.numel(),
                              _lib.stream_ptr(dev))
    _lib.check(rc, 'ty_flipflop_logz')
    _lib.count_launches(2)
    return trans


def _flipflop_make_trans_autograd(scores):
    """The same through autograd of `layers.flipflop_logpartition` (kept for the tests)."""
    x = scores.detach().float().requires_grad_()
    with torch.enable_grad():
        logz = layers.flipflop_logpartition(x).sum()
        trans, = torch.autograd.grad(logz, x)
    return trans.detach()
