#!/usr/bin/env python
"""Measured distance of the loss kernels from the fp64 oracle, next to the
distance of the REFERENCE'S OWN fp32 C (oracle/_ref) from the same fp64 oracle,
at the BASELINE sizes.  Writes a markdown table (profiles/r2_parity_table.md).

Why: north_star asks for 1e-4 relative fp32 against the reference.  Beyond
~1000 blocks the reference's fp32 recursion is itself further than that from
exact arithmetic on the small posterior entries, so "within 1e-4 of the
reference" is only meaningful above a floor set by the reference's own
round-off; this table is where the floors in tests/test_gpu_parity_sizes.py
come from.  Units: gradient errors are in units of a row's mass (|dG| * nblk;
a row of -nblk*grad is a posterior and sums to one).

    python tools/parity_table.py > profiles/r2_parity_table.md      (needs a GPU)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from taiyaki_b200 import ctc, layers  # noqa: E402

OFF = np.array([0, 1, 3, 4, 5], dtype=np.int32)
W = np.array([1.0, 1.0, 0.6, 1.0, 1.0], dtype=np.float32)
dev = torch.device('cuda:0')


def mods(raw, seed):
    rng = np.random.RandomState(seed)
    return np.concatenate([(r == 1).astype(np.int64) * rng.randint(0, 2, size=len(r)) for r in raw])


def stats(a, b, nblk):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64)) * nblk
    big = np.abs(b) * nblk > 1e-2          # entries carrying at least 1 % of a row's mass
    rel_big = (d[big] / (np.abs(b[big]) * nblk)).max() if big.any() else 0.0
    return d.max(), np.sqrt((d ** 2).mean()), rel_big


def row(name, nblk, ours, ref32, f64):
    o, r = stats(ours, f64, nblk), stats(ref32, f64, nblk)
    x = stats(ours, ref32, nblk)
    print('| %s | %.2e / %.2e / %.1e | %.2e / %.2e / %.1e | %.2e / %.2e / %.1e | %.2f |' % (
        name, o[0], o[1], o[2], r[0], r[1], r[2], x[0], x[1], x[2], o[0] / max(r[0], 1e-30)))


print('# Loss kernels vs fp64 oracle, next to the reference C (fp32) vs the same oracle\n')
print('Each cell: max |d| / rms |d| in units of a row\'s mass / max relative error over entries '
      'carrying >= 1 % of a row\'s mass.\n')
print('| case | GPU vs fp64 | reference C vs fp64 | GPU vs reference C | max ratio GPU/ref |')
print('|---|---|---|---|---|')

cases = [('A: crf nblk 800 N 64 S 40', 800, 64, 40, 5, None),
         ('crf nblk 2000 N 16 S 40', 2000, 16, 40, 2, None),
         ('B: cat-mod nblk 2000 N 64 S 45', 2000, 64, 45, 2, None),
         ('cat-mod nblk 2300 L 2048,300', 2300, 2, 45, 2, [2048, 300]),
         ('cat-mod nblk 1300 L 600..1', 1300, 6, 45, 2, [600, 513, 512, 511, 1, 40]),
         ('crf nblk 5000 L 4400,4000 (P=16)', 5000, 2, 40, 5, [4400, 4000])]
for name, nblk, nb, S, stride, lengths in cases:
    sc = oracle.synth_scores(nblk, nb, S, seed=nblk + nb)
    seqs, sl, raw = oracle.synth_seqs(nblk, nb, stride=stride, seed=nblk, lengths=lengths)
    x = torch.tensor(sc, device=dev, requires_grad=True)
    if S == 45:
        mc = mods(raw, nblk)
        c32, g32 = oracle.cat_mod_flipflop_loss(sc, seqs, sl, mc, OFF, W, 1.0, impl='ref')
        c64, g64 = oracle.cat_mod_flipflop_loss(sc, seqs, sl, mc, OFF, W, 1.0, impl='f64')
        cost = ctc.cat_mod_flipflop_loss(x, torch.tensor(seqs), torch.tensor(sl), torch.tensor(mc),
                                         OFF, W, 1.0)
    else:
        c32, g32 = oracle.crf_flipflop_loss(sc, seqs, sl, 1.0, impl='ref')
        c64, g64 = oracle.crf_flipflop_loss(sc, seqs, sl, 1.0, impl='f64')
        cost = ctc.crf_flipflop_loss(x, torch.tensor(seqs), torch.tensor(sl), 1.0)
    cost.sum().backward()
    row(name, nblk, x.grad.cpu().numpy(), g32, g64)
    print('| &nbsp;&nbsp;cost (relative) | %.1e | %.1e | %.1e | |' % (
        np.abs(cost.detach().cpu().numpy() / c64 - 1).max(), np.abs(c32 / c64 - 1).max(),
        np.abs(cost.detach().cpu().numpy() / c32 - 1).max()))

print('\n## Partition function (logZ) gradient vs fp64\n')
print('| case | GPU vs fp64: max |d| / rms / max rel over entries >= 1 % |')
print('|---|---|')
for nblk, nb in [(120, 4), (800, 64), (2000, 64)]:
    sc = oracle.synth_scores(nblk, nb, 40, seed=nblk)
    x = torch.tensor(sc, device=dev, requires_grad=True)
    layers.flipflop_logpartition(x).sum().backward()
    _, g64 = oracle.c_flipflop_logz(sc, want_grad=True, impl='f64')
    o = stats(x.grad.cpu().numpy(), g64, 1)
    print('| nblk %d N %d | %.2e / %.2e / %.1e |' % (nblk, nb, o[0], o[1], o[2]))
