#!/usr/bin/env python
"""Distances of the two CRF gradient paths from the fp64 restatement for tight lattices
(L close to nblk), in units of a row's mass."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_crf_fused as t  # noqa: E402

dev = torch.device('cuda:0')
for L in (129, 511, 512, 513, 1024, 1100):
    nblk = L + 40
    scores, seqs, seqlen, _ = t._inputs(nblk, 2, False, seed=L, lengths=[L, max(1, L // 3)])
    c1, g1, p1 = t._run(dev, scores, seqs, seqlen, None, 1.0, fused=True)
    c0, g0, p0 = t._run(dev, scores, seqs, seqlen, None, 1.0, fused=False)
    c64, g64 = t._oracle(scores, seqs, seqlen, None, 1.0)
    c32, g32 = __import__('oracle.oracle', fromlist=['x']).crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='ref')
    d = lambda g: (np.abs(g - g64).max() * nblk, np.sqrt(((g - g64) ** 2).mean()) * nblk)
    rel = lambda g: np.max(np.abs(g - g64) / (np.abs(g64) + 5e-6 / nblk))
    print('L %4d nblk %4d paths %d %d | fused max %.2e rms %.2e | pair max %.2e rms %.2e | ref C max %.2e rms %.2e | '
          'rel(floor 5e-6): fused %.2e pair %.2e refC %.2e | cost rel fused %.1e'
          % (L, nblk, p1, p0, *d(g1), *d(g0), *d(g32), rel(g1), rel(g0), rel(g32),
             np.max(np.abs(c1 - c64) / np.abs(c64))), flush=True)
