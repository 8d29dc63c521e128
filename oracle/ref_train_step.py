"""ref_train_step.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

CPU restatement of one optimiser step of the reference's
`bin/train_flipflop.py` on its stock code path, for bench.py's
`--impl reference` arm and `cpu_baseline`:

  * network: the reference's models/mLstm_flipflop.py / mGru_flipflop.py built
    from stock torch modules exactly as taiyaki/layers.py wraps them
    (Convolution :744-850 = ConstantPad1d + Conv1d + activation; Lstm :491-606
    = nn.LSTM with bias_hh frozen at 0; Reverse :117-153 = two flips;
    GlobalNormFlipFlop :1316-1411 = 5 tanh(Linear));
  * loss: ctc.pyx:116-153 (FlipFlopCRF) calling the REFERENCE's own C through
    oracle/_ref/libctc_ref.so (falls back to the restatement liboracle.so if
    the reference was never compiled), plus the TorchScript partition function
    layers.py:1253-1299 restated as the same torch ops;
  * step: zero_grad, backward, per-tensor max|grad| (apply_clipping
    :201-212), AdamW (train_flipflop.py:406-408).

Nothing here is imported by taiyaki_b200/.
"""
import time

import numpy as np
import torch
from torch import nn

from . import oracle


def swish(x):
    return x * torch.sigmoid(x)


class RefConvolution(nn.Module):
    def __init__(self, insize, size, winlen, stride=1, fun=torch.tanh):
        super().__init__()
        self.pad = nn.ConstantPad1d((winlen // 2, (winlen - 1) // 2), 0)
        self.conv = nn.Conv1d(insize, size, winlen, stride=stride)
        self.fun = fun

    def forward(self, x):
        return self.fun(self.conv(self.pad(x.permute(1, 2, 0)))).permute(2, 0, 1)


class RefRnn(nn.Module):
    def __init__(self, cell, size, reverse):
        super().__init__()
        self.rnn = (nn.LSTM if cell == 'lstm' else nn.GRU)(size, size)
        for name, p in self.rnn.named_parameters():
            if 'bias_hh' in name:
                p.requires_grad = False
                p.data.zero_()
        self.reverse = reverse

    def forward(self, x):
        if self.reverse:
            return torch.flip(self.rnn(torch.flip(x, (0,)))[0], (0,))
        return self.rnn(x)[0]


class RefFlipFlop(nn.Module):
    def __init__(self, size, nbase=4):
        super().__init__()
        self.linear = nn.Linear(size, 2 * nbase * (nbase + 1))

    def forward(self, x):
        return 5.0 * torch.tanh(self.linear(x))


def ref_network(cell='lstm', size=256, stride=None, winlen=19):
    """models/mLstm_flipflop.py:6-20 / models/mGru_flipflop.py:6-17"""
    if cell == 'lstm':
        stride = 5 if stride is None else stride
        front = [RefConvolution(1, 4, 5, 1, swish), RefConvolution(4, 16, 5, 1, swish),
                 RefConvolution(16, size, winlen, stride, swish)]
    else:
        stride = 2 if stride is None else stride
        front = [RefConvolution(1, size, winlen, stride, torch.tanh)]
    rnns = [RefRnn(cell, size, rev) for rev in (True, False, True, False, True)]
    return nn.Sequential(*front, *rnns, RefFlipFlop(size))


class RefFlipFlopCRF(torch.autograd.Function):
    """ctc.pyx:116-151 on the CPU through the reference's C library."""

    @staticmethod
    def forward(ctx, logprob, seqs, seqlen, sharpfact):
        impl = 'ref' if oracle.have_ref() else 'f32'
        cost, grad = oracle.crf_flipflop_loss(
            logprob.detach().numpy(), seqs.numpy(), seqlen.numpy(), sharpfact,
            want_grad=True, impl=impl)
        ctx.save_for_backward(torch.from_numpy(grad))
        return torch.from_numpy(cost)

    @staticmethod
    def backward(ctx, g):
        grad, = ctx.saved_tensors
        return grad * g.unsqueeze(1), None, None, None


def logaddexp(x, y):
    return torch.max(x, y) + nn.functional.softplus(-torch.abs(x - y))


def log_partition_flipflop(scores):
    """layers.py:1253-1299"""
    T, N, C = scores.shape
    nbase = oracle.nbase_flipflop(C)
    fwd = torch.cat([torch.zeros(N, nbase), torch.full((N, nbase), -50000.0)], 1)
    logZ = fwd.logsumexp(1, keepdim=True)
    fwd = fwd - logZ
    for scores_t in scores.unbind(0):
        curr = fwd.unsqueeze(1) + scores_t.reshape((-1, nbase + 1, 2 * nbase))
        base1 = curr[:, :nbase].logsumexp(2)
        base2 = logaddexp(curr[:, nbase, :nbase], curr[:, nbase, nbase:])
        new_state = torch.cat([base1, base2], dim=1)
        factors = new_state.logsumexp(1, keepdim=True)
        fwd = new_state - factors
        logZ = logZ + factors
    return logZ


class RefTrainer:
    def __init__(self, cell='lstm', size=256, seed=0, threads=None):
        if threads:
            torch.set_num_threads(threads)
        torch.manual_seed(seed)
        self.net = ref_network(cell, size)
        self.opt = torch.optim.AdamW(self.net.parameters(), lr=4e-3, betas=(0.9, 0.999),
                                     weight_decay=0.01, eps=1e-6)

    def step(self, indata, seqs, seqlens, sharpen=1.0):
        """indata [T_sig, N, 1] fp32 CPU; returns the scalar loss."""
        self.opt.zero_grad()
        out = self.net(indata)
        nblk = float(out.shape[0])
        lossvec = RefFlipFlopCRF.apply(out, seqs, seqlens, sharpen)
        lossvec = lossvec + log_partition_flipflop(out).squeeze(1) / nblk
        loss = lossvec.mean()
        loss.backward()
        fval = float(loss.detach())
        _ = [float(torch.max(torch.abs(p.grad))) for p in self.net.parameters()
             if p.requires_grad and p.grad is not None]
        self.opt.step()
        return fval


def time_reference(cell, t_sig, nchunks, steps, warmup, stride, seed=0, threads=None):
    """Samples/s of the CPU reference train step on a bounded sample."""
    tr = RefTrainer(cell, 256, seed, threads)
    nblk = -(-t_sig // stride)
    rng = np.random.RandomState(seed)
    x = torch.tensor(rng.standard_normal((t_sig, nchunks, 1)).astype(np.float32))
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nchunks, stride=stride, seed=seed + 1)
    seqs, seqlen = torch.tensor(seqs), torch.tensor(seqlen)
    for _ in range(warmup):
        tr.step(x, seqs, seqlen)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = tr.step(x, seqs, seqlen)
    dt = time.perf_counter() - t0
    return {'samples_per_s': steps * t_sig * nchunks / dt, 'ms_per_step': 1e3 * dt / steps,
            'loss': loss, 'threads': torch.get_num_threads(), 'nchunks': nchunks,
            'reference_c': oracle.have_ref()}
