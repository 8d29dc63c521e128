"""Layers on the flip-flop training path -- mirror of the relevant part of
taiyaki/layers.py (time-major [T, N, F] tensors, same class / attribute /
parameter names so `state_dict`s and model-definition files carry over).

  Lstm, GruMod              layers.py:491-725  recurrence in csrc/rnn_ws.cu / rnn_fp32.cu instead of cuDNN
  Reverse                   layers.py:117-153  loop direction instead of two flips
  Serial, Convolution       layers.py:944-982, :744-850
  GlobalNormFlipFlop[CatMod] layers.py:1316-1640
  flipflop_logpartition     layers.py:1875-1890 -> csrc/logz.cu (was CuPy / TorchScript)
  global_norm_flipflop, log_partition_flipflop  layers.py:1277-1313
"""
from collections import OrderedDict

import numpy as np
import torch
from scipy import linalg
from scipy.stats import truncnorm
from torch import nn

from . import _lib, activation, flipflopfings

MODEL_VERSION = 3
_CELL_LSTM, _CELL_GRU = 0, 1


# ---------------------------------------------------------------------------
# initialisers (layers.py:22-114)
def init_(param, value):
    value_as_tensor = torch.tensor(np.asarray(value), dtype=param.data.dtype)
    with torch.no_grad():
        param.set_(value_as_tensor.to(param.device))


def random_orthonormal(n, m=None):
    """QR of Gaussian noise with the sign fix of Mezzadri (layers.py:37-66)."""
    m = n if m is None else m
    assert m >= n
    x = np.random.randn(m, m)
    Q, r = linalg.qr(x, mode='economic')
    flipper = np.diag(np.sign(np.diag(r)))
    return Q.dot(flipper)[:n, :]


def orthonormal_matrix(nrow, ncol):
    """Block-orthonormal [nrow, ncol] matrix (layers.py:69-96)."""
    nrep = nrow // ncol
    out = np.zeros((nrow, ncol), dtype='f4')
    for i in range(nrep):
        out[i * ncol: i * ncol + ncol] = random_orthonormal(ncol)
    remsize = nrow - nrep * ncol
    if remsize > 0:
        out[nrep * ncol:, :] = random_orthonormal(remsize, ncol)
    return out


def truncated_normal(size, sd):
    """Normal truncated at +/-2 sd (layers.py:99-114)."""
    res = sd * truncnorm.rvs(-2, 2, size=size)
    return res.astype('f4')


def _reshape(x, shape):
    return x.reshape(shape)


#: Operand type of the dense projection GEMMs around the recurrence
#: (x W_ih^T, and the weight / input gradients): 'bf16' (tensor cores, fp32
#: accumulate and fp32 result -- north_star's "dense bf16 contractions"),
#: 'tf32' (what cuDNN does for the reference on Ampere and later) or 'fp32'
#: (the parity mode: plain fp32 products, see `set_precision`).
PROJECTION_DTYPE = 'bf16'


class _matmul_tf32:
    def __init__(self, allow):
        self.allow = allow

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.allow

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def _mm(a, b):
    """a @ b with fp32 result through the library: bf16 operands with fp32
    accumulation; fp32 operands as TF32 or, in the parity mode, true fp32."""
    if a.dtype == torch.bfloat16:
        return torch.mm(a, b, out_dtype=torch.float32)
    with _matmul_tf32(PROJECTION_DTYPE != 'fp32'):
        return torch.mm(a, b)


def set_precision(mode):
    """'bf16' (default): bf16 tensor-core products everywhere north_star allows them --
    tcgen05 GEMMs for the projections, the bf16 recurrent product of csrc/rnn_ws.cu.
    'fp32': the parity mode -- every product of the recurrent layers, the strided
    convolution and the score projection in fp32 (csrc/rnn_fp32.cu for the recurrence),
    which reproduces torch.nn.LSTM / nn.GRU fp32 to ~1e-5; ~10x slower."""
    global PROJECTION_DTYPE, RNN_IMPL
    assert mode in ('bf16', 'fp32')
    PROJECTION_DTYPE = mode
    RNN_IMPL = 'ws' if mode == 'bf16' else 'fp32'


def _operand(t):
    return t.to(torch.bfloat16) if PROJECTION_DTYPE == 'bf16' else t


#: bf16 contractions go to csrc/gemm_tc5.cu (TMA + tcgen05); False sends them to the
#: library GEMM instead (A/B timing only -- `TY_GEMM=lib` in the environment)
import os as _os
TC5_GEMM = _os.environ.get('TY_GEMM', 'tc5') != 'lib'
_EPI_STORE, _EPI_BIAS_TANH, _EPI_ATOMIC, _EPI_REDUCE = 0, 1, 2, 3


def _tc5_ok(*operands):
    """bf16, unit inner stride, 16-byte aligned base and row pitch."""
    return TC5_GEMM and all(
        t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and
        t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0 for t in operands)


def _gemm(a, a_mn, b, b_mn, M, N, K, out=None, epi=_EPI_STORE, bias=None, scale=1.0,
          k_splits=1, map_g=0, map_h=0):
    """out[M, N] (fp32) = A[M, K] B[N, K]^T on the tensor cores (ty_gemm_bf16).
    a_mn / b_mn: the operand tensor is the stored TRANSPOSE ([K, M] / [K, N])."""
    if out is None:
        # row pitch a multiple of 16 bytes (TMA store); a [:, :N] view when N % 4 != 0
        out = torch.empty(M, (N + 3) // 4 * 4, dtype=torch.float32, device=a.device)[:, :N]
    rc = _lib.lib().ty_gemm_bf16(
        _lib.ptr(a), a.stride(0), int(a_mn), _lib.ptr(b), b.stride(0), int(b_mn), M, N, K,
        _lib.ptr(out), out.stride(0), epi, _lib.ptr(bias), float(scale), int(k_splits),
        int(map_g), int(map_h), _lib.stream_ptr(a.device))
    _lib.check(rc, 'ty_gemm_bf16')
    _lib.count_launches(1)
    return out


def _k_splits(M, N, K, ctas=296):
    """CTAs sharing an output tile of a weight-gradient product (contraction over
    time x batch): enough of them to fill the chip twice over."""
    tiles = -(-M // 128) * -(-N // (64 if N <= 64 else 128))
    return max(1, min(-(-K // 64), ctas // tiles))


def _mm_nt(x, w):
    """x[M, K] w[N, K]^T -> fp32 [M, N]"""
    if _tc5_ok(x, w):
        return _gemm(x, 0, w, 0, x.shape[0], w.shape[0], x.shape[1])
    return _mm(x, w.t())


def _mm_nn(d, w):
    """d[M, K] w[K, N] -> fp32 [M, N] (input gradient: w read in place, MN-major)"""
    if _tc5_ok(d, w):
        return _gemm(d, 0, w, 1, d.shape[0], w.shape[1], d.shape[1])
    return _mm(d, w)


def _mm_tn(d, x, out=None, map_g=0, map_h=0):
    """d[K, M]^T x[K, N] -> fp32 [M, N] (weight gradient: both operands read in
    place, MN-major; split-K, accumulated into `out` (zeros when None), rows
    optionally sent from unit-major to gate-major order)."""
    K, M = d.shape
    N = x.shape[1]
    if _tc5_ok(d, x):
        if out is None:
            out = torch.zeros(M, N, dtype=torch.float32, device=d.device)
        return _gemm(d, 1, x, 1, M, N, K, out=out, epi=_EPI_ATOMIC, k_splits=_k_splits(M, N, K),
                     map_g=map_g, map_h=map_h)
    res = _mm(d.t(), x)
    if map_g:
        res = _gate_major(res, map_g)
    if out is not None:
        return out.add_(res)
    return res


# ---------------------------------------------------------------------------
# recurrence
def _unit_major(w, G):
    """Row permutation gate-major [G*H, K] -> unit-major [H*G, K] (row u*G+g <- g*H+u)."""
    GH, K = w.shape
    return w.view(G, GH // G, K).transpose(0, 1).reshape(GH, K)


def _gate_major(w, G):
    """Inverse of `_unit_major`."""
    GH, K = w.shape
    return w.view(GH // G, G, K).transpose(0, 1).reshape(GH, K)


#: When True (set by training.TrainStep around loss.backward()), the weight-gradient
#: GEMMs of the recurrent layers run on a side stream, concurrently with the next
#: layer's backward recurrence (which occupies 64 of the 148 SMs), and are added into
#: `param.grad` there; `flush_weight_grads()` joins the side stream.  Off by default:
#: plain autograd semantics (gradients returned through the graph).
DEFER_WEIGHT_GRADS = False
_pending = []          # (event, tensors kept alive until the join)
_side_streams = {}
#: called as hook(last_parameter_of_the_layer, side_stream) at the start of a recurrent
#: layer's backward, when every gradient of the layers ABOVE it is final (their weight-
#: gradient GEMMs are on the side stream, everything else was accumulated on the current
#: stream): training.FlatGradients uses it to start the all-reduce of those slices early
GRADS_FINAL_ABOVE_HOOK = None


def _side_stream(device):
    key = device.index
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def flush_weight_grads():
    """Make the current stream wait for every deferred weight-gradient update."""
    global _pending
    for ev, _keep in _pending:
        torch.cuda.current_stream().wait_event(ev)
    _pending = []


#: 'ws' = warp-specialised bf16 tensor-core kernels behind the unit-major ABI
#: (csrc/rnn_ws.cuh; hidden sizes CLUSTER_KERNEL_SIZES, which include the reference's
#: defaults 256 / 384 and the 96 of its "fast" models); 'fp32' = the fp32
#: recurrence behind the gate-major ABI (csrc/rnn_fp32.cu; any hidden size) -- selected
#: by `set_precision('fp32')`, and automatically for sizes the cluster kernels lack
RNN_IMPL = 'ws'
CLUSTER_KERNEL_SIZES = (32, 64, 96, 128, 160, 192, 224, 256, 320, 384, 448)   # = ty_rnn_um_supported


_warned_sizes = set()


def _warn_slow_size(H):
    if H not in _warned_sizes:
        _warned_sizes.add(H)
        import warnings
        warnings.warn('taiyaki_b200: hidden size %d has no bf16 cluster kernel (sizes: %s); the recurrence '
                      'runs through the fp32 parity kernels, ~50x slower' % (H, list(CLUSTER_KERNEL_SIZES)))


class _Recurrence(torch.autograd.Function):
    """y = RNN(x) with the sequential part in csrc/rnn_ws.cu (or csrc/rnn_fp32.cu).

    forward:  xproj = x W_ih^T (one GEMM) -> ty_rnn_forward_* (adds b_ih)
    backward: ty_rnn_backward_* gives d xproj, the bias gradient (and the
              hidden-side gradient for the GRU); weight / input gradients are
              dense GEMMs over all time steps.
    With the unit-major kernels W_ih is row-permuted while it is converted to
    the GEMM operand type, so the projection comes out as [T, N, H, G]; the
    weight gradients are permuted back.
    """

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, cell, reverse, x16):
        _lib.require_cuda(x, 'x')
        lib = _lib.lib()
        T, N, I = x.shape
        G = 4 if cell == _CELL_LSTM else 3
        H = w_hh.shape[1]
        use16 = PROJECTION_DTYPE == 'bf16'
        um = use16 and RNN_IMPL == 'ws' and H in CLUSTER_KERNEL_SIZES
        if use16 and RNN_IMPL == 'ws' and not um:
            _warn_slow_size(H)
        w_hh_c = w_hh.detach().contiguous().float()
        if use16 and x16 is not None:
            xo = x16.view(T * N, I)            # bf16 copy written by the producing layer
        else:
            xo = _operand(x.contiguous().float().view(T * N, I))
        if um:     # permute the rows while converting: one small copy kernel
            wo = torch.empty(G * H, I, dtype=torch.bfloat16, device=x.device)
            wo.view(H, G, I).copy_(w_ih.detach().view(G, H, I).transpose(0, 1))
        else:
            wo = _operand(w_ih.detach())
        xproj = _mm_nt(xo, wo)
        bias = b_ih.detach().contiguous().float() if b_ih is not None else None
        dev = x.device
        y = torch.empty(T, N, H, dtype=torch.float32, device=dev)
        y16 = torch.empty(T, N, H, dtype=torch.bfloat16, device=dev) if use16 else None
        reserve = torch.empty(lib.ty_rnn_reserve_bytes(cell, T, N, H) // 4,
                              dtype=torch.float32, device=dev)
        fwd = lib.ty_rnn_forward_um if um else lib.ty_rnn_forward_ex
        with _lib.timed('rnn_fwd', dev):
            rc = fwd(cell, _lib.ptr(xproj), _lib.ptr(bias), _lib.ptr(w_hh_c),
                     T, N, H, int(reverse), _lib.ptr(y), _lib.ptr(y16),
                     _lib.ptr(reserve), _lib.stream_ptr(dev))
        _lib.check(rc, 'ty_rnn_forward')
        _lib.count_launches(1)
        ctx.save_for_backward(xo, wo, w_hh_c, y, reserve, y16)
        ctx.cfg = (cell, bool(reverse), b_ih is not None, G, H, I, um)
        ctx.weights = (w_ih, w_hh)
        ctx.last_param = b_ih if b_ih is not None else w_hh
        if y16 is None:
            y16 = y.new_empty(0)
        ctx.mark_non_differentiable(y16)
        ctx.set_materialize_grads(False)     # no zero tensor for the bf16 side output
        return y, y16

    @staticmethod
    def backward(ctx, dy, _unused):
        lib = _lib.lib()
        xo, wo, w_hh_c, y, reserve, y16 = ctx.saved_tensors
        cell, reverse, has_bias, G, H, I, um = ctx.cfg
        T, N, _ = y.shape
        dev = y.device
        if GRADS_FINAL_ABOVE_HOOK is not None:
            GRADS_FINAL_ABOVE_HOOK(ctx.last_param, _side_stream(dev))
        use16 = y16 is not None
        gdt = torch.bfloat16 if use16 else torch.float32
        if dy is None:
            dy = torch.zeros_like(y)
        dy = dy.contiguous().float()
        # gradients w.r.t. the projections are only ever GEMM operands: the kernel
        # writes them in the operand type; the bias gradient is summed in fp32 inside
        do = torch.empty(T, N, G * H, dtype=gdt, device=dev)
        db = torch.zeros(G * H, dtype=torch.float32, device=dev) if has_bias else None
        yo = y16 if use16 else y
        # h_{t-1} of every step is y shifted by one step along the loop direction
        sl_cur, sl_prev = (slice(None, -1), slice(1, None)) if reverse else \
            (slice(1, None), slice(None, -1))
        hp2 = yo[sl_prev].reshape(-1, H)
        if um:
            dhid = torch.empty(T, N, G * H, dtype=gdt, device=dev) if cell == _CELL_GRU else None
            with _lib.timed('rnn_bwd', dev):
                rc = lib.ty_rnn_backward_um(cell, _lib.ptr(dy), _lib.ptr(w_hh_c), T, N, H,
                                            int(reverse), _lib.ptr(y), _lib.ptr(reserve),
                                            _lib.ptr(do), _lib.ptr(dhid), _lib.ptr(db),
                                            _lib.stream_ptr(dev))
            _lib.check(rc, 'ty_rnn_backward_um')
            _lib.count_launches(1)
            d2 = do.view(T * N, G * H)
            dh_side = do if cell == _CELL_LSTM else dhid
            w_ih, w_hh = ctx.weights
            dh2 = dh_side[sl_cur].reshape(-1, G * H)
            if DEFER_WEIGHT_GRADS and w_ih.is_leaf and w_hh.is_leaf:
                # straight into param.grad (a view of the flat gradient buffer): split-K
                # partial products are added with red.global.add, rows permuted back to the
                # parameters' gate-major order on the way out
                for w in (w_ih, w_hh):
                    if w.grad is None:
                        w.grad = torch.zeros_like(w)
                main = torch.cuda.current_stream(dev)
                side = _side_stream(dev)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    if ctx.needs_input_grad[1]:
                        _mm_tn(d2, xo, out=w_ih.grad, map_g=G, map_h=H)
                    if ctx.needs_input_grad[2]:
                        _mm_tn(dh2, hp2, out=w_hh.grad, map_g=G, map_h=H)
                    done = torch.cuda.Event()
                    done.record(side)
                _pending.append((done, (do, dhid, xo, yo)))
                dx = _mm_nn(d2, wo).view(T, N, I) if ctx.needs_input_grad[0] else None
                return dx, None, None, db, None, None, None
            dx = _mm_nn(d2, wo).view(T, N, I) if ctx.needs_input_grad[0] else None
            dw_ih = _mm_tn(d2, xo, map_g=G, map_h=H)
            dw_hh = _mm_tn(dh2, hp2, map_g=G, map_h=H)
            return dx, dw_ih, dw_hh, db, None, None, None
        dhn = torch.empty(T, N, H, dtype=gdt, device=dev) if cell == _CELL_GRU else None
        with _lib.timed('rnn_bwd', dev):
            rc = lib.ty_rnn_backward_ex(cell, _lib.ptr(dy), _lib.ptr(w_hh_c), T, N, H,
                                        int(reverse), _lib.ptr(y), _lib.ptr(reserve),
                                        _lib.ptr(do), _lib.ptr(dhn), int(use16), _lib.ptr(db),
                                        _lib.stream_ptr(dev))
        _lib.check(rc, 'ty_rnn_backward_ex')
        _lib.count_launches(1)
        d2 = do.view(T * N, G * H)
        d_cur = do[sl_cur]
        dx = _mm_nn(d2, wo).view(T, N, I) if ctx.needs_input_grad[0] else None
        dw_ih = _mm_tn(d2, xo)
        if cell == _CELL_LSTM:
            dw_hh = _mm_tn(d_cur.reshape(-1, G * H), hp2)
        else:
            dw_hh = torch.cat([
                _mm_tn(d_cur[:, :, :2 * H].reshape(-1, 2 * H), hp2),
                _mm_tn(dhn[sl_cur].reshape(-1, H), hp2)], 0)
        return dx, dw_ih, dw_hh, db, None, None, None


def _run_recurrence(x, w_ih, w_hh, b_ih, cell, reverse):
    """Apply the recurrence; the bf16 copy of the output rides along on the
    tensor so the next recurrent layer can use it as its GEMM operand."""
    x16 = getattr(x, '_ty_bf16', None)
    if x16 is not None and (x16.shape != x.shape or x16.device != x.device):
        x16 = None
    y, y16 = _Recurrence.apply(x, w_ih, w_hh, b_ih, cell, reverse, x16)
    if y16.numel():
        y._ty_bf16 = y16
    return y


class Lstm(nn.Module):
    """LSTM layer (layers.py:491-606).  `self.lstm` is kept as the parameter
    container so checkpoints keep `lstm.weight_ih_l0` etc.; `bias_hh` stays
    frozen at zero and hidden from `named_parameters`."""

    def __init__(self, insize, size, has_bias=True):
        super().__init__()
        self.lstm = nn.LSTM(insize, size, bias=has_bias)
        self.insize = insize
        self.size = size
        self.has_bias = has_bias
        self._disable_state_bias()
        self.reset_parameters()

    def _disable_state_bias(self):
        for name, param in self.lstm.named_parameters():
            if 'bias_hh' in name:
                param.requires_grad = False
                param.data.zero_()

    def reset_parameters(self):
        for name, param in self.named_parameters():
            shape = list(param.shape)
            if 'weight_hh' in name or 'weight_ih' in name:
                init_(param, orthonormal_matrix(*shape))
            else:
                init_(param, truncated_normal(shape, sd=0.5))

    def named_parameters(self, prefix='', recurse=True):
        for name, param in self.lstm.named_parameters(prefix=prefix, recurse=recurse):
            if 'bias_hh' not in name:
                yield name, param

    def forward(self, x, reverse=False):
        m = self.lstm
        return _run_recurrence(x, m.weight_ih_l0, m.weight_hh_l0,
                               m.bias_ih_l0 if self.has_bias else None, _CELL_LSTM, reverse)

    def json(self):
        res = OrderedDict([('type', "LSTM"), ('activation', "tanh"), ('gate', "sigmoid"),
                           ('size', self.size), ('insize', self.insize),
                           ('bias', self.has_bias)])
        res['params'] = OrderedDict([
            ('iW', _reshape(self.lstm.weight_ih_l0, (4, self.size, self.insize))),
            ('sW', _reshape(self.lstm.weight_hh_l0, (4, self.size, self.size))),
            ('b', _reshape(self.lstm.bias_ih_l0, (4, self.size)))])
        return res


class GruMod(nn.Module):
    """Guppy-compatible GRU (layers.py:609-725): cuDNN 'linear before reset'
    form with the hidden bias frozen at zero."""

    def __init__(self, insize, size, has_bias=True):
        super().__init__()
        self.cudnn_gru = nn.GRU(insize, size, bias=has_bias)
        self.insize = insize
        self.size = size
        self.has_bias = has_bias
        self._disable_state_bias()
        self.reset_parameters()

    def reset_parameters(self):
        for name, param in self.named_parameters():
            shape = list(param.shape)
            if 'weight_hh' in name or 'weight_ih' in name:
                init_(param, orthonormal_matrix(*shape))
            else:
                init_(param, truncated_normal(shape, sd=0.5))

    def _disable_state_bias(self):
        for name, param in self.cudnn_gru.named_parameters():
            if 'bias_hh' in name:
                param.requires_grad = False
                param.data.zero_()

    def named_parameters(self, prefix='', recurse=True):
        prefix = prefix + ('.' if prefix else '')
        for name, param in self.cudnn_gru.named_parameters(recurse=recurse):
            if 'bias_hh' not in name:
                yield prefix + name, param

    def forward(self, x, reverse=False):
        m = self.cudnn_gru
        return _run_recurrence(x, m.weight_ih_l0, m.weight_hh_l0,
                               m.bias_ih_l0 if self.has_bias else None, _CELL_GRU, reverse)

    def json(self):
        res = OrderedDict([('type', "GruMod"), ('activation', "tanh"), ('gate', "sigmoid"),
                           ('size', self.size), ('insize', self.insize),
                           ('bias', self.has_bias)])
        iW = _cudnn_to_guppy_gru(self.cudnn_gru.weight_ih_l0)
        sW = _cudnn_to_guppy_gru(self.cudnn_gru.weight_hh_l0)
        b = _cudnn_to_guppy_gru(self.cudnn_gru.bias_ih_l0)
        res['params'] = OrderedDict([
            ('iW', _reshape(iW, (3, self.size, self.insize))),
            ('sW', _reshape(sW, (3, self.size, self.size))),
            ('b', _reshape(b, (3, self.size)))])
        return res


def _cudnn_to_guppy_gru(p):
    """cuDNN (r, z, n) -> Guppy (z, r, n) ordering (layers.py:728-741)."""
    x, y, z = torch.chunk(p, 3)
    return torch.cat([y, x, z], 0)


class Reverse(nn.Module):
    """Run the enclosed layer backwards in time (layers.py:117-153).  The
    recurrent layers of this package take the direction as a flag, which saves
    the two flipped copies of the activation tensor; any other layer is wrapped
    between two flips as in the reference."""

    def __init__(self, layer):
        super().__init__()
        self.layer = layer

    def forward(self, x):
        if isinstance(self.layer, (Lstm, GruMod)):
            return self.layer(x, reverse=True)
        return torch.flip(self.layer(torch.flip(x, (0,))), (0,))

    def json(self):
        return OrderedDict([('type', "reverse"), ('sublayers', self.layer.json())])


class Serial(nn.Module):
    """Apply layers one after the other (layers.py:944-982)."""

    def __init__(self, layers):
        super().__init__()
        self.sublayers = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.sublayers:
            x = layer(x)
        return x

    def json(self):
        return OrderedDict([('type', "serial"),
                            ('sublayers', [layer.json() for layer in self.sublayers])])


class _ConvTimeMajor(torch.autograd.Function):
    """1D convolution of a time-major [T, N, C] tensor as window-gather + one
    dense GEMM, producing time-major [T_out, N, C_out] directly (no TBF<->BFT
    permutes, layers.py:816-831).  Same arithmetic as nn.Conv1d on the
    zero-padded signal.  Wide layers (C*k >= 64, the strided feature layer):
    the gather writes the bf16 GEMM operand directly (ty_im2col_time_major_bf16)
    with a column of ones, so the bias is part of the GEMM and its gradient a
    column of the weight-gradient GEMM.  Narrow layers: fp32 / TF32 operands."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding):
        T, N, C = x.shape
        Cout, _, k = weight.shape
        Tp = T + padding[0] + padding[1]
        Tout = (Tp - k) // stride + 1
        wide = C * k >= 64 and x.is_cuda and PROJECTION_DTYPE == 'bf16'
        if wide:
            ld = (C * k + 1 + 7) // 8 * 8
            xc = x.detach().contiguous().float()
            co = torch.empty(Tout * N, ld, dtype=torch.bfloat16, device=x.device)
            rc = _lib.lib().ty_im2col_time_major_bf16(
                _lib.ptr(xc), T, N, C, k, stride, padding[0], Tout, ld, _lib.ptr(co),
                _lib.stream_ptr(x.device))
            _lib.check(rc, 'ty_im2col_time_major_bf16')
            _lib.count_launches(1)
            wo = torch.zeros(Cout, ld, dtype=torch.bfloat16, device=x.device)
            wo[:, :C * k] = weight.detach().reshape(Cout, C * k)
            if bias is not None:
                wo[:, C * k] = bias.detach()
            out = _mm_nt(co, wo)
        else:
            ld = C * k
            xp = torch.nn.functional.pad(x, (0, 0, 0, 0, padding[0], padding[1]))
            co = xp.unfold(0, k, stride).reshape(Tout * N, C * k)     # [T_out*N, C*k]
            wo = weight.detach().reshape(Cout, C * k)
            out = _mm(co, wo.t())
            if bias is not None:
                out += bias.detach()
        ctx.save_for_backward(co, wo)
        ctx.cfg = (T, N, C, Cout, k, stride, padding, Tp, Tout, wide, bias is not None, ld)
        return out.view(Tout, N, Cout)

    @staticmethod
    def backward(ctx, dout):
        co, wo = ctx.saved_tensors
        T, N, C, Cout, k, stride, padding, Tp, Tout, wide, has_bias, ld = ctx.cfg
        d2 = dout.reshape(Tout * N, Cout)
        do = _operand(d2) if wide else d2
        dwf = _mm_tn(do, co) if wide else _mm(do.t(), co)            # [Cout, ld]
        if wide:
            dw = dwf[:, :C * k].reshape(Cout, C, k)
            db = dwf[:, C * k] if has_bias else None
        else:
            dw = dwf.view(Cout, C, k)
            db = d2.sum(0) if has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            dcols = _mm_nn(do, wo) if wide else _mm(do, wo)         # [T_out*N, ld]
            if dcols.is_cuda:
                dx = torch.empty(T, N, C, dtype=torch.float32, device=dcols.device)
                rc = _lib.lib().ty_col2im_time_major_ld(
                    _lib.ptr(dcols), ld, Tout, N, C, k, stride, padding[0], T, _lib.ptr(dx),
                    _lib.stream_ptr(dcols.device))
                _lib.check(rc, 'ty_col2im_time_major_ld')
                _lib.count_launches(1)
            else:
                dcols = dcols.view(Tout, N, C * k).permute(1, 2, 0)     # [N, C*k, T_out]
                dxp = torch.nn.functional.fold(dcols, output_size=(1, Tp), kernel_size=(1, k),
                                               stride=(1, stride))      # [N, C, 1, Tp]
                dx = dxp[:, :, 0, padding[0]:padding[0] + T].permute(2, 0, 1)
        return dx, dw, db, None, None


_ACT_CODES = {activation.linear: 0, activation.tanh: 1, activation.swish: 2}


class _ConvSmall(torch.autograd.Function):
    """Stride-1 convolution with few channels + activation as direct kernels
    (csrc/conv.cu: conv_small_*): one launch forward, three backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, pad_left, act):
        T, N, C = x.shape
        Cout, _, k = weight.shape
        xc = x.detach().contiguous().float()
        wc = weight.detach().contiguous().float()
        bc = bias.detach().contiguous().float() if bias is not None else None
        z = torch.empty(T, N, Cout, dtype=torch.float32, device=x.device)
        a = torch.empty_like(z)
        rc = _lib.lib().ty_conv_small_forward(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc), T, N, C,
                                              Cout, k, pad_left, act, _lib.ptr(z), _lib.ptr(a),
                                              _lib.stream_ptr(x.device))
        _lib.check(rc, 'ty_conv_small_forward')
        _lib.count_launches(1)
        ctx.save_for_backward(xc, wc, z)
        ctx.cfg = (pad_left, act, bias is not None)
        return a

    @staticmethod
    def backward(ctx, da):
        xc, wc, z = ctx.saved_tensors
        pad_left, act, has_bias = ctx.cfg
        T, N, C = xc.shape
        Cout, _, k = wc.shape
        da = da.contiguous().float()
        dz = torch.empty_like(z)
        dw = torch.zeros_like(wc)
        db = torch.zeros(Cout, dtype=torch.float32, device=xc.device) if has_bias else None
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        rc = _lib.lib().ty_conv_small_backward(
            _lib.ptr(da), _lib.ptr(z), _lib.ptr(xc), _lib.ptr(wc), T, N, C, Cout, k, pad_left,
            act, _lib.ptr(dz), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(dx),
            _lib.stream_ptr(xc.device))
        _lib.check(rc, 'ty_conv_small_backward')
        _lib.count_launches(3 if dx is not None else 2)
        return dx, dw, db, None, None


class _ConvIn1(torch.autograd.Function):
    """Strided convolution of the single-channel signal (first layer of the mGru models,
    Convolution(1, size, 19, stride=2)) as direct fp32 kernels (csrc/conv.cu: conv_in1_*): as a
    GEMM it is too narrow for a tensor-core tile and bf16 would round the raw signal.  Returns the
    pre-activation; the signal carries no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding):
        T, N, _ = x.shape
        Cout, _, k = weight.shape
        Tout = (T + padding[0] + padding[1] - k) // stride + 1
        xc = x.detach().contiguous().float()
        wc = weight.detach().contiguous().float()
        bc = bias.detach().contiguous().float() if bias is not None else None
        z = torch.empty(Tout, N, Cout, dtype=torch.float32, device=x.device)
        rc = _lib.lib().ty_conv_in1_forward(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc), T, N, Cout, k,
                                            stride, padding[0], Tout, _lib.ptr(z),
                                            _lib.stream_ptr(x.device))
        _lib.check(rc, 'ty_conv_in1_forward')
        _lib.count_launches(1)
        ctx.save_for_backward(xc)
        ctx.cfg = (T, N, Cout, k, stride, padding[0], Tout, bias is not None)
        return z

    @staticmethod
    def backward(ctx, dz):
        xc, = ctx.saved_tensors
        T, N, Cout, k, stride, pad_left, Tout, has_bias = ctx.cfg
        dz = dz.contiguous().float()
        dw = torch.zeros(Cout, 1, k, dtype=torch.float32, device=dz.device)
        db = torch.zeros(Cout, dtype=torch.float32, device=dz.device) if has_bias else None
        rc = _lib.lib().ty_conv_in1_wgrad(_lib.ptr(xc), _lib.ptr(dz), T, N, Cout, k, stride, pad_left,
                                          Tout, _lib.ptr(dw), _lib.ptr(db), _lib.stream_ptr(dz.device))
        _lib.check(rc, 'ty_conv_in1_wgrad')
        _lib.count_launches(1)
        return None, dw, db, None, None


class Convolution(nn.Module):
    """1D convolution over time for [T, N, F] tensors (layers.py:744-850)."""

    def __init__(self, insize, size, winlen, stride=1, pad=None, fun=activation.tanh,
                 has_bias=True):
        super().__init__()
        self.has_bias = has_bias
        self.insize = insize
        self.size = size
        self.stride = stride
        self.winlen = winlen
        if pad is None:
            pad = (winlen // 2, (winlen - 1) // 2)
        self.padding = pad
        self.pad = nn.ConstantPad1d(pad, 0)
        self.conv = nn.Conv1d(kernel_size=winlen, in_channels=insize, out_channels=size,
                              stride=stride, bias=has_bias)
        self.activation = fun
        self.reset_parameters()

    def reset_parameters(self):
        winit = orthonormal_matrix(self.conv.weight.shape[0],
                                   int(np.prod(self.conv.weight.shape[1:])))
        init_(self.conv.weight, winit.reshape(self.conv.weight.shape))
        if self.has_bias:
            init_(self.conv.bias, truncated_normal(list(self.conv.bias.shape), sd=0.5))

    def forward(self, x):
        if x.is_cuda:
            act = _ACT_CODES.get(self.activation, None)
            if (self.stride == 1 and act is not None and x.dim() == 3 and
                    self.padding[0] + self.padding[1] == self.winlen - 1 and
                    _lib.lib().ty_conv_small_supported(self.insize, self.size, self.winlen)):
                return _ConvSmall.apply(x, self.conv.weight, self.conv.bias, self.padding[0], act)
            if (self.insize == 1 and x.dim() == 3 and not x.requires_grad and
                    _lib.lib().ty_conv_in1_supported(1, self.size, self.winlen)):
                out = _ConvIn1.apply(x, self.conv.weight, self.conv.bias, self.stride, self.padding)
            else:
                out = _ConvTimeMajor.apply(x, self.conv.weight, self.conv.bias, self.stride,
                                           self.padding)
            return self.activation(out)
        x = x.permute(1, 2, 0)
        out = self.activation(self.conv(self.pad(x)))
        return out.permute(2, 0, 1)

    def json(self):
        res = OrderedDict([("type", "convolution"), ("insize", self.insize),
                           ("size", self.size), ("bias", self.has_bias),
                           ("winlen", self.conv.kernel_size[0]),
                           ("stride", self.conv.stride[0]), ("padding", self.padding),
                           ("activation", self.activation.__name__)])
        res['params'] = OrderedDict(
            [("W", self.conv.weight)] + [("b", self.conv.bias)] if self.has_bias else [])
        return res


class _ProjLinear(torch.autograd.Function):
    """y = x W^T + b as a dense bf16 contraction with fp32 accumulation and fp32
    result (the same operand policy as the recurrent layers' projections); the
    bf16 copy of x written by a preceding recurrent layer is used when present."""

    @staticmethod
    def forward(ctx, x, weight, bias, x16):
        T, N, I = x.shape
        xo = x16.view(T * N, I) if x16 is not None else _operand(x.reshape(T * N, I))
        wo = _operand(weight.detach())
        out = _mm_nt(xo, wo)
        if bias is not None:
            out += bias.detach()
        ctx.save_for_backward(xo, wo)
        ctx.has_bias = bias is not None
        return out.view(T, N, -1)

    @staticmethod
    def backward(ctx, dy):
        xo, wo = ctx.saved_tensors
        T, N, O = dy.shape
        d2 = dy.reshape(T * N, O)
        do = _operand(d2)
        dx = _mm_nn(do, wo).view(T, N, -1) if ctx.needs_input_grad[0] else None
        dw = _mm_tn(do, xo)
        db = d2.sum(0) if ctx.has_bias else None
        return dx, dw, db, None


class _ScoreProjection(torch.autograd.Function):
    """scores = scale * tanh(x W^T + b) as ONE kernel: the bf16 contraction on the
    tensor cores with bias, tanh and the scale applied to the fp32 accumulators
    before they leave the SM (layers.py:1402-1411 is a cuBLAS GEMM followed by
    three elementwise passes over [T, N, S])."""

    @staticmethod
    def forward(ctx, x, weight, bias, x16, scale):
        T, N, I = x.shape
        O = weight.shape[0]
        xo = x16.view(T * N, I) if x16 is not None else _operand(x.reshape(T * N, I))
        wo = _operand(weight.detach())
        b = bias.detach().float().contiguous() if bias is not None else None
        out = _gemm(xo, 0, wo, 0, T * N, O, I, epi=_EPI_BIAS_TANH, bias=b, scale=scale)
        ctx.save_for_backward(xo, wo, out)
        ctx.cfg = (bias is not None, float(scale))
        return out.view(T, N, O)

    @staticmethod
    def backward(ctx, dy):
        xo, wo, out = ctx.saved_tensors
        has_bias, scale = ctx.cfg
        T, N, O = dy.shape
        # d/dz scale tanh(z) = scale - s^2 / scale with s the saved output
        dz = dy.reshape(T * N, O) * (scale - out * out * (1.0 / scale))
        do = dz.to(torch.bfloat16)
        dx = _mm_nn(do, wo).view(T, N, -1) if ctx.needs_input_grad[0] else None
        dw = _mm_tn(do, xo)
        db = dz.sum(0) if has_bias else None
        return dx, dw, db, None, None


def _proj_linear(linear, x):
    """`linear(x)` for a time-major CUDA tensor through the bf16 projection path."""
    if not x.is_cuda or x.dim() != 3 or PROJECTION_DTYPE != 'bf16':
        return linear(x)
    x16 = getattr(x, '_ty_bf16', None)
    if x16 is not None and (x16.shape != x.shape or x16.device != x.device):
        x16 = None
    return _ProjLinear.apply(x, linear.weight, linear.bias, x16)


class GlobalNormFlipFlop(nn.Module):
    """scale * fun(x W + b) transition scores (layers.py:1316-1411); global
    normalisation is the loss function's job."""

    def __init__(self, insize, nbase, has_bias=True, fun=activation.tanh, scale=5.0):
        super().__init__()
        self.insize = insize
        self.nbase = nbase
        self.size = flipflopfings.nstate_flipflop(nbase)
        self.activation = fun
        self.has_bias = has_bias
        self.linear = nn.Linear(insize, self.size, bias=has_bias)
        self.reset_parameters()
        self.scale = scale

    def json(self):
        res = OrderedDict([('type', 'GlobalNormTwoState'), ('size', self.size),
                           ('insize', self.insize), ('bias', self.has_bias),
                           ('scale', self.scale), ("activation", self.activation.__name__)])
        res['params'] = OrderedDict(
            [('W', self.linear.weight)] + [('b', self.linear.bias)] if self.has_bias else [])
        return res

    def reset_parameters(self):
        init_(self.linear.weight, orthonormal_matrix(*list(self.linear.weight.shape)))
        if self.has_bias:
            init_(self.linear.bias, truncated_normal(list(self.linear.bias.shape), sd=0.5))

    def forward(self, x):
        if (self.activation is activation.tanh and x.is_cuda and x.dim() == 3 and
                PROJECTION_DTYPE == 'bf16' and TC5_GEMM and self.insize % 8 == 0):
            x16 = getattr(x, '_ty_bf16', None)
            if x16 is not None and (x16.shape != x.shape or x16.device != x.device):
                x16 = None
            return _ScoreProjection.apply(x, self.linear.weight, self.linear.bias, x16,
                                          self.scale)
        return self.scale * self.activation(_proj_linear(self.linear, x))


class GlobalNormFlipFlopCatMod(nn.Module):
    """Flip-flop transition scores plus a categorical modified-base stream
    (layers.py:1414-1640).  Output columns: 2*ncan*(ncan+1) transition scores,
    then per-canonical-base log-softmax groups (canonical first)."""

    def compute_label_conversions(self):
        can_labels, mod_labels = [], []
        can_grouped_mods = dict((can_b, 0) for can_b in self.can_bases)
        for b, can_b in zip(self.alphabet, self.collapse_alphabet):
            can_labels.append(self.can_bases.find(can_b))
            if b in self.can_bases:
                mod_labels.append(0)
            else:
                can_grouped_mods[can_b] += 1
                mod_labels.append(can_grouped_mods[can_b])
        self.can_labels = np.array(can_labels)
        self.mod_labels = np.array(mod_labels)

    def compute_layer_mods_info(self):
        self.output_alphabet = ''
        for can_b in self.can_bases:
            self.output_alphabet += can_b
            for b, can_bi in zip(self.alphabet, self.collapse_alphabet):
                if can_bi == can_b and b != can_b:
                    self.output_alphabet += b
        self.ordered_mod_long_names = (
            None if self.mod_long_names is None else
            [self.mod_name_conv[b] for b in self.alphabet if b in self.mod_bases])
        self.can_nmods = np.array([
            sum(b == can_b for b in self.collapse_alphabet) - 1 for can_b in self.can_bases])
        self.can_mods_offsets = np.cumsum(np.concatenate(
            [[0], self.can_nmods + 1])).astype(np.int32)
        self.can_indices = []
        curr_n_mods = 0
        for bi_nmods in self.can_nmods:
            self.can_indices.append(np.concatenate([
                [0], np.arange(curr_n_mods + 1, curr_n_mods + 1 + bi_nmods)]))
            curr_n_mods += bi_nmods

    def __init__(self, insize, alphabet_info, has_bias=True):
        super().__init__()
        self.insize = insize
        self.has_bias = has_bias
        self.alphabet = alphabet_info.alphabet
        self.collapse_alphabet = alphabet_info.collapse_alphabet
        self.mod_long_names = alphabet_info.mod_long_names
        self.mod_name_conv = alphabet_info.mod_name_conv
        self.can_bases = alphabet_info.can_bases
        self.mod_bases = alphabet_info.mod_bases
        self.ncan_base = alphabet_info.ncan_base
        self.nmod_base = alphabet_info.nmod_base
        self.compute_label_conversions()
        self.compute_layer_mods_info()
        self.ntrans_states = 2 * self.ncan_base * (self.ncan_base + 1)
        self.size = self.ntrans_states + 1 + self.nmod_base
        self.lsm = nn.LogSoftmax(2)
        self.linear = nn.Linear(insize, self.size, bias=self.has_bias)
        self.reset_parameters()

    @property
    def nbase(self):
        return self.ncan_base

    def json(self):
        res = OrderedDict([('type', 'GlobalNormTwoStateCatMod'), ('size', self.size),
                           ('insize', self.insize), ('bias', self.has_bias),
                           ('can_nmods', self.can_nmods),
                           ('output_alphabet', self.output_alphabet),
                           ('modified_base_long_names', self.ordered_mod_long_names)])
        res['params'] = OrderedDict(
            [('W', self.linear.weight)] + [('b', self.linear.bias)] if self.has_bias else [])
        return res

    def reset_parameters(self):
        init_(self.linear.weight, orthonormal_matrix(*list(self.linear.weight.shape)))
        if self.has_bias:
            init_(self.linear.bias, truncated_normal(list(self.linear.bias.shape), sd=0.5))

    def get_softmax_cat_mods(self, cat_mod_scores):
        # index tensors live on the scores' device (indexing with the numpy arrays copies them to
        # the device on every call, which a CUDA-graph capture does not even allow)
        # (kept outside the module so that they are not pickled into checkpoints)
        key = (tuple(tuple(int(i) for i in ix) for ix in self.can_indices), cat_mod_scores.device)
        if key not in _CAN_INDEX_CACHE:
            _CAN_INDEX_CACHE[key] = [torch.as_tensor(np.asarray(ix), dtype=torch.long, device=key[1])
                                     for ix in self.can_indices]
        mod_layers = []
        for lab_indices in _CAN_INDEX_CACHE[key]:
            mod_layers.append(self.lsm(cat_mod_scores.index_select(2, lab_indices)))
        return torch.cat(mod_layers, dim=2)

    def forward(self, x):
        y = _proj_linear(self.linear, x)
        trans_scores = 5.0 * activation.tanh(y[:, :, :self.ntrans_states])
        cat_mod_scores = y[:, :, self.ntrans_states:]
        assert cat_mod_scores.shape[2] == self.nmod_base + 1, (
            'Invalid scores provided to forward:  Expected: {}  got: {}'.format(
                self.nmod_base + 1, cat_mod_scores.shape[2]))
        cat_mod_scores = self.get_softmax_cat_mods(cat_mod_scores)
        assert cat_mod_scores.shape[2] == self.nmod_base + self.ncan_base, (
            'Invalid softmax categorical mod scores:  Expected: {}  got: {}'.format(
                self.nmod_base + self.ncan_base, cat_mod_scores.shape[2]))
        return torch.cat((trans_scores, cat_mod_scores), dim=2)


_CAN_INDEX_CACHE = {}


def is_cat_mod_model(net):
    """layers.py:1643-1656"""
    assert isinstance(net, Serial)
    return isinstance(net.sublayers[-1], GlobalNormFlipFlopCatMod)


# ---------------------------------------------------------------------------
# partition function
class LogZ(torch.autograd.Function):
    """log partition function of the flip-flop CRF and its gradient, one pass
    over the scores (replaces cupy_extensions/flipflop.py:338-354)."""

    @staticmethod
    def forward(ctx, scores):
        _lib.require_cuda(scores, 'scores')
        lib = _lib.lib()
        T, N, S = scores.shape
        nbase = flipflopfings.nbase_flipflop(S)
        x = scores.detach()
        if x.dtype != torch.float32:
            x = x.float()
        # a [:, :, :S] view of a wider contiguous tensor is used in place
        if x.stride(2) == 1 and x.stride(0) == N * x.stride(1) and x.stride(1) >= S:
            ld = x.stride(1)
        else:
            x = x.contiguous()
            ld = S
        device = x.device
        logz = torch.empty(N, dtype=torch.float32, device=device)
        want_grad = scores.requires_grad
        grad = torch.empty(T, N, S, dtype=torch.float32, device=device) if want_grad else None
        ws = _lib.workspace(lib.ty_flipflop_logz_workspace_bytes(nbase, T, N), device)
        rc = lib.ty_flipflop_logz(_lib.ptr(x), ld, T, N, nbase, 1.0, _lib.ptr(logz), 1.0,
                                  _lib.ptr(grad), S, 0, _lib.ptr(ws), ws.numel(),
                                  _lib.stream_ptr(device))
        _lib.check(rc, 'ty_flipflop_logz')
        _lib.count_launches(2 if want_grad else 1)
        if want_grad:
            ctx.save_for_backward(grad)
        return logz

    @staticmethod
    def backward(ctx, g):
        grad, = ctx.saved_tensors
        return grad * g[:, None]


def flipflop_logpartition(x, _never_use_cupy=False):
    """Log-partition function of the flip-flop model, [T, N, S] -> [N]
    (layers.py:1875-1890).  `_never_use_cupy` is accepted for signature
    compatibility; there is a single device implementation."""
    return LogZ.apply(x)


def log_partition_flipflop(scores):
    """[T, N, S] -> [N, 1] (layers.py:1277-1299)"""
    return LogZ.apply(scores).unsqueeze(1)


def global_norm_flipflop(scores):
    """scores - logZ / T (layers.py:1302-1313)"""
    T = scores.shape[0]
    return scores - log_partition_flipflop(scores) / np.float32(T)
