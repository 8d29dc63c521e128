#!/usr/bin/env python
"""Forward time of one Lstm(256, 256) layer at config A (T 800, N 64) with the library named by
TY_B200_LIB: A/B timing of patched copies of csrc/rnn_ws.cuh (ablations: reserve stores off, x ring
loads off; see profiles/r1_rnn_stalls.md)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from taiyaki_b200 import layers
dev = torch.device('cuda:0')
torch.manual_seed(0); np.random.seed(0)
mod = layers.Lstm(256, 256).to(dev)
x = torch.randn(800, 64, 256, device=dev)
def run():
    with torch.no_grad():
        mod(x)
for _ in range(5): run()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(os.environ.get('TY_B200_LIB'), 'layer fwd (GEMM + recurrence) median %.4f ms min %.4f' % (np.median(ts), min(ts)))
