#!/usr/bin/env python
"""Short workload for ncu: each hot kernel a few times at BASELINE config A
(nblk=800, N=64, S=40, H=256).  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 \
      -o gpurun_out/prof python tools/profile_target.py [crf|logz|rnn|gru|aux|remap]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from taiyaki_b200 import ctc, layers  # noqa: E402

dev = torch.device('cuda:0')
which = sys.argv[1:] or ['crf', 'logz', 'rnn']
nblk, N = 800, 64
scores = torch.tensor(oracle.synth_scores(nblk, N, 40, seed=0), device=dev)
seqs, seqlen, _ = oracle.synth_seqs(nblk, N, stride=5, seed=1)
reps = int(os.environ.get('TY_PROF_REPS', '2'))
for _ in range(reps):
    if 'crf' in which:
        ctc.crf_flipflop_cost_grad(scores, torch.tensor(seqs), torch.tensor(seqlen), 1.0, True)
    if 'logz' in which:
        layers.flipflop_logpartition(scores.detach().requires_grad_(True))
    if 'aux' in which:      # batching, decode and convolution kernels
        from taiyaki_b200 import chunk_selection, decode, device_batching, signal_mapping, training
        from taiyaki_b200.activation import swish
        reads = signal_mapping.synthetic_reads(24, seed=7)
        fp = chunk_selection.sample_filter_parameters(reads, 100, 4000, 10.0, 10.0, 0.1, 5, 1.1)
        store = device_batching.DeviceReadStore(reads, dev)
        md = training.NETWORK_METADATA(False, True, False)
        store.sample(64, 4000, fp, md, 4)
        decode.flipflop_viterbi(scores)
        decode.flipflop_make_trans(scores)
        net = layers.Serial([layers.Convolution(1, 4, 5, fun=swish), layers.Convolution(4, 16, 5, fun=swish),
                             layers.Convolution(16, 256, 19, stride=5, fun=swish)]).to(dev)
        xin = torch.randn(4000, N, 1, device=dev, requires_grad=True)
        net(xin).sum().backward()
    if 'remap' in which:    # remapping DP: 148 reads of 8000 blocks x 3500 positions, one launch
        from taiyaki_b200 import flipflop_remap
        rng = np.random.RandomState(1)
        rseqs = [''.join('ACGT'[b] for b in rng.randint(0, 4, size=3500)) for _ in range(148)]
        rscores = [torch.randn(8000, 40, device=dev) for _ in range(148)]
        flipflop_remap.flipflop_remap_batch(rscores, rseqs)
    if 'gru' in which:
        torch.manual_seed(0)
        np.random.seed(0)
        mod = layers.GruMod(256, 256).to(dev)
        x = torch.randn(nblk, N, 256, device=dev, requires_grad=True)
        y = mod(x)
        y.backward(torch.ones_like(y))
    if 'rnn' in which:
        torch.manual_seed(0)
        np.random.seed(0)
        mod = layers.Lstm(256, 256).to(dev)
        x = torch.randn(nblk, N, 256, device=dev, requires_grad=True)
        y = mod(x)
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print('profile target done')
