#!/usr/bin/env python
"""Kernel-level timings on one GPU (CUDA events, warm-up, L2 flush between
iterations).  Writes JSON lines to gpurun_out/microbench.jsonl.  Not the
contract benchmark (that is bench.py); this is the tuning loop."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (synthetic input generators only)
from taiyaki_b200 import _lib, ctc, layers  # noqa: E402

dev = torch.device('cuda:0')
OUT = os.path.join(ROOT, 'gpurun_out', 'microbench.jsonl')
os.makedirs(os.path.dirname(OUT), exist_ok=True)
_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return -1


def spin(ms=300):
    """Keep the GPU busy so the clocks leave the idle state before timing."""
    a = torch.randn(4096, 4096, device=dev)
    t0 = time.time()
    while (time.time() - t0) * 1e3 < ms:
        (a @ a).sum().item()


def timeit(fn, iters=10, warmup=3):
    spin()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def emit(**kw):
    kw['sm_mhz'] = sm_clock()
    print(json.dumps(kw))
    with open(OUT, 'a') as f:
        f.write(json.dumps(kw) + '\n')


def bench_crf(nblk, nbatch, ntrans, stride, tag):
    scores = torch.tensor(oracle.synth_scores(nblk, nbatch, ntrans, seed=0), device=dev)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, stride=stride, seed=1)
    seqs_t, seqlen_t = torch.tensor(seqs), torch.tensor(seqlen)
    alg_bytes = 2 * ntrans * 4 * nblk * nbatch
    if ntrans > 40:
        mod_cats = torch.tensor(np.concatenate(
            [((r == 1) & (np.random.RandomState(3).uniform(size=len(r)) < 0.5)).astype(np.int64)
             for r in raw]))
        off = np.array([0, 1, 3, 4, 5], dtype=np.int32)
        w = np.ones(5, dtype=np.float32)

        def run(want_grad):
            xx = scores.detach().requires_grad_(want_grad)
            return ctc.cat_mod_flipflop_loss(xx, seqs_t, seqlen_t, mod_cats, off, w, 1.0)
    else:
        def run(want_grad):
            return ctc.crf_flipflop_cost_grad(scores, seqs_t, seqlen_t, 1.0, want_grad)
    from taiyaki_b200 import _lib
    lib = _lib.lib()
    for fused, P in [(1, 0), (0, 0), (0, 1), (0, 2), (0, 4), (0, 8)]:
        if nblk > 2000 and P in (1, 2):
            continue
        lib.ty_crf_tuning(P, fused)
        try:
            med, mn = timeit(lambda: run(True))
        except Exception as e:
            emit(what='crf_grad', tag=tag, P=P, fused=fused, error=str(e))
            continue
        emit(what='crf_grad(indices+chain+post)', tag=tag, nblk=nblk, N=nbatch, S=ntrans, P=P or 'auto',
             fused=fused, ms_median=med, ms_min=mn, alg_GBps=alg_bytes / med / 1e6,
             mean_L=float(np.mean(seqlen)))
    lib.ty_crf_tuning(0, 1)
    med, mn = timeit(lambda: run(False))
    emit(what='crf_cost_only', tag=tag, nblk=nblk, N=nbatch, ms_median=med, ms_min=mn)
    x = scores.detach()[:, :, :40]
    xg = x.detach().clone().requires_grad_(True)
    med, mn = timeit(lambda: layers.flipflop_logpartition(xg))
    emit(what='logz+posterior', tag=tag, nblk=nblk, N=nbatch, ms_median=med, ms_min=mn)
    med, mn = timeit(lambda: layers.flipflop_logpartition(x))
    emit(what='logz_only', tag=tag, nblk=nblk, N=nbatch, ms_median=med, ms_min=mn)
    if oracle.libcupy_ref() is not None:
        # the reference's GPU path: its CuPy RawKernels compiled with nvcc (oracle/build_cupy_ref.py)
        med, mn = timeit(lambda: oracle.cupy_ref_logz(x))
        emit(what='reference_gpu_logz+posterior (CuPy RawKernels under nvcc)', tag=tag, nblk=nblk, N=nbatch,
             ms_median=med, ms_min=mn)
        med, mn = timeit(lambda: oracle.cupy_ref_logz(x, want_trans=False))
        emit(what='reference_gpu_logz_only (CuPy RawKernels under nvcc)', tag=tag, nblk=nblk, N=nbatch,
             ms_median=med, ms_min=mn)


def bench_rnn(cell, T, N, H, tag, kernels_only=False):
    torch.manual_seed(0)
    if kernels_only:
        dy = torch.randn(T, N, H, device=dev)
        return bench_rnn_kernels(cell, T, N, H, tag, dy)
    mod = (layers.Lstm(H, H) if cell == 'lstm' else layers.GruMod(H, H)).to(dev)
    nnmod = (torch.nn.LSTM(H, H) if cell == 'lstm' else torch.nn.GRU(H, H)).to(dev)
    x = torch.randn(T, N, H, device=dev, requires_grad=True)
    dy = torch.randn(T, N, H, device=dev)

    def ours_fwd():
        with torch.no_grad():
            mod(x)

    def ours_fb():
        y = mod(x)
        y.backward(dy)

    def cudnn_fwd():
        with torch.no_grad():
            nnmod(x)

    def cudnn_fb():
        y = nnmod(x)[0]
        y.backward(dy)

    for name, fn in [('ours_fwd', ours_fwd), ('ours_fwd+bwd', ours_fb),
                     ('cudnn_fwd', cudnn_fwd), ('cudnn_fwd+bwd', cudnn_fb)]:
        try:
            med, mn = timeit(fn, iters=5, warmup=2)
            emit(what='rnn_layer', cell=cell, impl=name, tag=tag, T=T, N=N, H=H, ms_median=med,
                 ms_min=mn, us_per_step=1e3 * med / T)
        except Exception as e:
            emit(what='rnn_layer', cell=cell, impl=name, tag=tag, error=str(e)[:300])
    bench_rnn_kernels(cell, T, N, H, tag, dy)


def bench_rnn_kernels(cell, T, N, H, tag, dy):
    # recurrence kernel alone (no projections)
    lib = _lib.lib()
    G = 4 if cell == 'lstm' else 3
    code = 0 if cell == 'lstm' else 1
    xproj = torch.randn(T, N, G * H, device=dev)
    w_hh = torch.randn(G * H, H, device=dev) / np.sqrt(H)
    y = torch.empty(T, N, H, device=dev)
    reserve = torch.empty(lib.ty_rnn_reserve_bytes(code, T, N, H) // 4, device=dev)
    dxp = torch.empty(T, N, G * H, device=dev)
    dhn = torch.empty(T, N, H, device=dev)
    st = _lib.stream_ptr(dev)
    if cell == 'lstm':
        f = lambda: lib.ty_lstm_forward(_lib.ptr(xproj), None, _lib.ptr(w_hh), T, N, H, 0, _lib.ptr(y), _lib.ptr(reserve), st)
        b = lambda: lib.ty_lstm_backward(_lib.ptr(dy), _lib.ptr(w_hh), T, N, H, 0, _lib.ptr(y), _lib.ptr(reserve), _lib.ptr(dxp), None, st)
    else:
        f = lambda: lib.ty_gru_forward(_lib.ptr(xproj), None, _lib.ptr(w_hh), T, N, H, 0, _lib.ptr(y), _lib.ptr(reserve), st)
        b = lambda: lib.ty_gru_backward(_lib.ptr(dy), _lib.ptr(w_hh), T, N, H, 0, _lib.ptr(y), _lib.ptr(reserve), _lib.ptr(dxp), _lib.ptr(dhn), None, st)
    dx16 = torch.empty(T, N, G * H, device=dev, dtype=torch.bfloat16)
    dh16 = torch.empty(T, N, G * H, device=dev, dtype=torch.bfloat16)
    y16 = torch.empty(T, N, H, device=dev, dtype=torch.bfloat16)
    fu = lambda: lib.ty_rnn_forward_um(code, _lib.ptr(xproj), None, _lib.ptr(w_hh), T, N, H, 0, _lib.ptr(y), _lib.ptr(y16), _lib.ptr(reserve), st)
    bu = lambda: lib.ty_rnn_backward_um(code, _lib.ptr(dy), _lib.ptr(w_hh), T, N, H, 0, _lib.ptr(y), _lib.ptr(reserve), _lib.ptr(dx16), _lib.ptr(dh16), None, st)
    for name, fn in [('kernel_fwd', f), ('kernel_bwd', b), ('kernel_fwd_um', fu), ('kernel_bwd_um', bu)]:
        med, mn = timeit(fn, iters=5, warmup=2)
        emit(what='rnn_kernel', cell=cell, impl=name, tag=tag, T=T, N=N, H=H, ms_median=med,
             ms_min=mn, us_per_step=1e3 * med / T, variant=os.environ.get('TY_RNN_FWD', ''),
             y_checksum=float(y.double().abs().sum()))


if __name__ == '__main__':
    which = sys.argv[1:] or ['crf', 'rnn']
    t0 = time.time()
    if 'crf' in which:
        bench_crf(800, 64, 40, 5, 'A')
        bench_crf(2000, 64, 45, 2, 'B')
    if 'rnn' in which:
        bench_rnn('lstm', 800, 64, 256, 'A')
        bench_rnn('gru', 2000, 64, 256, 'B')
    if 'rnnk' in which:
        bench_rnn('lstm', 800, 64, 256, 'A', kernels_only=True)
        bench_rnn('gru', 2000, 64, 256, 'B', kernels_only=True)
    if 'sweep' in which:
        for nblk in (1000, 2000, 4000, 8000):
            bench_crf(nblk, 64, 40, 5, 'E%d' % nblk)
    emit(what='done', seconds=time.time() - t0)
