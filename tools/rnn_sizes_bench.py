#!/usr/bin/env python
"""Recurrent layer (Lstm / GruMod module, forward and forward+backward) per hidden size at
T=800, N=64: the cluster kernels of csrc/rnn_ws.cuh against cuDNN (torch.nn) on the same GPU.
Usage under gpurun: python tools/rnn_sizes_bench.py > gpurun_out/rnn_sizes.jsonl"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiyaki_b200 import layers  # noqa: E402

dev = torch.device('cuda:0')
T, N = 800, 64


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


sizes = [int(v) for v in sys.argv[1:]] or [96, 256, 384, 448, 288]
for cell in ('lstm', 'gru'):
    for H in sizes:
        torch.manual_seed(0)
        np.random.seed(0)
        mod = (layers.Lstm(H, H) if cell == 'lstm' else layers.GruMod(H, H)).to(dev)
        nnmod = (torch.nn.LSTM(H, H) if cell == 'lstm' else torch.nn.GRU(H, H)).to(dev)
        x = torch.randn(T, N, H, device=dev, requires_grad=True)
        dy = torch.randn(T, N, H, device=dev)

        def fwd(m=mod):
            with torch.no_grad():
                return m(x)

        def fb(m=mod):
            y = m(x)
            y = y[0] if isinstance(y, tuple) else y
            y.backward(dy)

        rec = {'what': 'rnn_layer_by_size', 'cell': cell, 'H': H, 'T': T, 'N': N,
               'cluster_kernel': H in layers.CLUSTER_KERNEL_SIZES,
               'ours_fwd_ms': timeit(fwd), 'ours_fwd_bwd_ms': timeit(fb),
               'cudnn_fwd_ms': timeit(lambda: fwd(nnmod)), 'cudnn_fwd_bwd_ms': timeit(lambda: fb(nnmod))}
        print(json.dumps(rec), flush=True)
