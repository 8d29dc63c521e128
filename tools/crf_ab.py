#!/usr/bin/env python
"""A/B of the two CRF gradient paths on the same inputs: chains with the posterior fused in
(csrc/crf_fused.cu) against chain kernel + posterior kernel (csrc/crf_flipflop.cu); prints the
timing of each and the largest difference between their gradients.
Usage under gpurun: [TY_B200_LIB=build_variants/lib_x.so] python tools/crf_ab.py [A B ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from taiyaki_b200 import _lib, ctc  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.lib()
CONFIGS = {'A': (800, 64, 40, 5), 'B': (2000, 64, 45, 5), 'T1000': (200, 64, 40, 5), 'T2000': (400, 64, 40, 5),
           'T8000': (1600, 64, 40, 5)}


def timeit(fn, reps=20):
    for _ in range(3):
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
    return float(np.median(ts)), float(np.min(ts))


for tag in (sys.argv[1:] or ['A', 'B']):
    nblk, nbatch, ntrans, stride = CONFIGS[tag]
    scores = torch.tensor(oracle.synth_scores(nblk, nbatch, ntrans, seed=0), device=dev)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, stride=stride, seed=1)
    seqs_t, seqlen_t = torch.tensor(seqs), torch.tensor(seqlen)
    if ntrans > 40:
        mod_cats = torch.tensor(np.concatenate(
            [((r == 1) & (np.random.RandomState(3).uniform(size=len(r)) < 0.5)).astype(np.int64) for r in raw]))
        off = np.array([0, 1, 3, 4, 5], dtype=np.int32)
        w = np.ones(5, dtype=np.float32)

        def run():
            xx = scores.detach().requires_grad_(True)
            loss = ctc.cat_mod_flipflop_loss(xx, seqs_t, seqlen_t, mod_cats, off, w, 1.0)
            loss.sum().backward()
            return loss.detach(), xx.grad
    else:
        def run():
            return ctc.crf_flipflop_cost_grad(scores, seqs_t, seqlen_t, 1.0, True)
    out = {}
    for fused in (1, 0):
        lib.ty_crf_tuning(0, fused)
        med, mn = timeit(run)
        cost, grad = run()
        out[fused] = (cost.clone(), grad.clone())
        print(json.dumps({'what': 'crf_grad', 'tag': tag, 'nblk': nblk, 'N': nbatch, 'S': ntrans, 'fused': fused,
                          'ms_median': med, 'ms_min': mn, 'path': lib.ty_crf_last_path(), 'lib': os.path.basename(_lib.LIB_PATH)}), flush=True)
    lib.ty_crf_tuning(0, 1)
    dc = (out[1][0] - out[0][0]).abs().max().item()
    dg = (out[1][1] - out[0][1]).abs().max().item()
    print(json.dumps({'what': 'fused_vs_two_kernel', 'tag': tag, 'max_abs_diff_cost': dc, 'max_abs_diff_grad': dg,
                      'grad_abs_max': out[0][1].abs().max().item()}), flush=True)
