#!/usr/bin/env python
"""Remapping DP (csrc/remap.cu) timing: NREAD reads of T blocks x L positions in one launch
against the numpy restatement of the reference loop on one read.  One JSON line each."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from taiyaki_b200 import _lib, flipflop_remap  # noqa: E402

dev = torch.device('cuda:0')
for nread, T, L in ((1, 8000, 3500), (148, 8000, 3500), (148, 2000, 900), (592, 2000, 900)):
    rng = np.random.RandomState(1)
    seqs = [''.join('ACGT'[b] for b in rng.randint(0, 4, size=L)) for _ in range(nread)]
    scores = [torch.randn(T, 40, device=dev) for _ in range(nread)]
    flipflop_remap.flipflop_remap_batch(scores, seqs)
    torch.cuda.synchronize()
    t0 = time.time()
    res = flipflop_remap.flipflop_remap_batch(scores, seqs)
    torch.cuda.synchronize()
    dt = time.time() - t0
    _lib.PROFILE = {}
    flipflop_remap.flipflop_remap_batch(scores, seqs)
    torch.cuda.synchronize()
    kernel_ms = _lib.PROFILE['remap'][0][0].elapsed_time(_lib.PROFILE['remap'][0][1])
    _lib.PROFILE = None
    step, stay = flipflop_remap.remap_indices(seqs[0])
    s0 = scores[0].cpu().numpy()
    t0 = time.time()
    oscore, opath = oracle.map_to_crf_viterbi(s0, step, stay)
    cpu = time.time() - t0
    assert res[0][0] == oscore and (res[0][1] == opath).all()
    print(json.dumps({'what': 'flipflop_remap', 'reads': nread, 'T': T, 'L': L,
                      'gpu_ms_per_launch_incl_host': round(dt * 1e3, 2),
                      'kernel_ms': round(kernel_ms, 3), 'kernel_ms_per_read': round(kernel_ms / nread, 4),
                      'cpu_numpy_ms_per_read': round(cpu * 1e3, 1),
                      'cells_per_s': round(nread * T * L / kernel_ms / 1e6, 2), 'unit': 'G cells/s'}), flush=True)
