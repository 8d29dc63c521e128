#!/usr/bin/env python
"""Fused training loss (ty_flipflop_train_loss through ctc.flipflop_train_loss) at the
BASELINE shapes: config A (S=40, nblk=800) and B (cat-mod S=45, nblk=2000), CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (synthetic input generators only)
from taiyaki_b200 import ctc  # noqa: E402

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(tag, nblk, N, S, stride):
    scores = torch.tensor(oracle.synth_scores(nblk, N, S, seed=0), device=dev).requires_grad_(True)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, N, stride=stride, seed=1)
    seqs_t, seqlen_t = torch.tensor(seqs).to(dev), torch.tensor(seqlen).to(dev)
    ctc.hint_lengths(seqlen_t, int(seqlen.max()), int(seqlen.sum()))
    kw = {}
    if S > 40:
        mod_cats = torch.tensor(np.concatenate(
            [((r == 1) & (np.random.RandomState(3).uniform(size=len(r)) < 0.5)).astype(np.int64)
             for r in raw])).to(dev)
        kw = dict(mod_cats=mod_cats, can_mods_offsets=np.array([0, 1, 3, 4, 5], dtype=np.int32),
                  mod_cat_weights=np.ones(5, dtype=np.float32))
    ts = []
    for i in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctc.flipflop_train_loss(scores, seqs_t, seqlen_t, 1.0, **kw)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    alg = 2 * S * 4 * nblk * N
    print(json.dumps({'what': 'fused_train_loss', 'tag': tag, 'nblk': nblk, 'N': N, 'S': S,
                      'serial': os.environ.get('TY_LOSS_SERIAL', '0'),
                      'ms_median': float(np.median(ts)), 'ms_min': float(np.min(ts)),
                      'alg_GBps': alg / np.median(ts) / 1e6}))


run('A', 800, 64, 40, 5)
run('B', 2000, 64, 45, 2)
run('B40', 2000, 64, 40, 2)
