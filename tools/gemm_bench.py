#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM (csrc/gemm_tc5.cu) against the library GEMM
(torch.mm -> cuBLAS nvjet) on the contractions of one train step at config A.
CUDA events around each launch, L2 flushed between iterations (256 MB write),
median of 20.  Prints one JSON line per shape."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taiyaki_b200 import layers  # noqa: E402

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    ts = []
    for _ in range(n + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[3:]))


def report(name, M, N, K, t_tc5, t_lib, out_bytes, in_bytes):
    flops = 2.0 * M * N * K
    print(json.dumps({
        'shape': name, 'M': M, 'N': N, 'K': K, 'tc5_us': round(t_tc5 * 1e3, 1),
        'lib_us': round(t_lib * 1e3, 1), 'tc5_tflops': round(flops / t_tc5 / 1e9, 1),
        'tc5_gbs': round((out_bytes + in_bytes) / t_tc5 / 1e6, 0),
        'lib_gbs': round((out_bytes + in_bytes) / t_lib / 1e6, 0)}), flush=True)


T, Nb, H, G = 800, 64, 256, 4
M = T * Nb
x = torch.randn(M, H, device=dev).to(torch.bfloat16)
w = torch.randn(G * H, H, device=dev).to(torch.bfloat16)
d = torch.randn(M, G * H, device=dev).to(torch.bfloat16)
wt = w.t().contiguous()

layers.TC5_GEMM = True
t1 = timeit(lambda: layers._mm_nt(x, w))
t2 = timeit(lambda: torch.mm(x, w.t(), out_dtype=torch.float32))
report('input projection x W_ih^T', M, G * H, H, t1, t2, M * G * H * 4, M * H * 2)

t1 = timeit(lambda: layers._mm_nn(d, w))
t2 = timeit(lambda: torch.mm(d, w, out_dtype=torch.float32))
report('input gradient dG W_ih', M, H, G * H, t1, t2, M * H * 4, M * G * H * 2)

acc = torch.zeros(G * H, H, device=dev)
t1 = timeit(lambda: layers._mm_tn(d, x, out=acc, map_g=G, map_h=H))
t2 = timeit(lambda: torch.mm(d.t(), x, out_dtype=torch.float32))
report('weight gradient dG^T X (split-K, red.add)', G * H, H, M, t1, t2, G * H * H * 4, M * (G * H + H) * 2)

ws = torch.randn(40, H, device=dev).to(torch.bfloat16)
bias = torch.randn(40, device=dev)
t1 = timeit(lambda: layers._gemm(x, 0, ws, 0, M, 40, H, epi=1, bias=bias, scale=5.0))
t2 = timeit(lambda: 5.0 * torch.tanh(torch.mm(x, ws.t(), out_dtype=torch.float32) + bias))
report('score projection + bias + 5 tanh', M, 40, H, t1, t2, M * 40 * 4, M * H * 2)

co = torch.randn(M, 312, device=dev).to(torch.bfloat16)
wc = torch.randn(H, 312, device=dev).to(torch.bfloat16)
t1 = timeit(lambda: layers._mm_nt(co, wc))
t2 = timeit(lambda: torch.mm(co, wc.t(), out_dtype=torch.float32))
report('strided convolution (im2col) x W^T', M, H, 312, t1, t2, M * H * 4, M * 312 * 2)

# ---- per-CTA timeline of the input projection (globaltimer stamps, ns) ----
import ctypes
from taiyaki_b200 import _lib
lib = _lib.lib()
lib.ty_gemm_debug_timeline.argtypes = [ctypes.c_void_p]
lib.ty_gemm_debug_timeline.restype = None
for name, fn, ncta in []:
    buf = torch.zeros(ncta * 8, dtype=torch.int64, device=dev)
    lib.ty_gemm_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
    flush.zero_()
    fn()
    torch.cuda.synchronize()
    lib.ty_gemm_debug_timeline(None)
    t = buf.cpu().numpy().reshape(ncta, 8).astype(np.float64)
    t0 = t[:, 0].min()
    rel = t - t[:, :1]
    names = ['start', 'setup done', 'first operands landed', 'last MMA issued', 'accumulator ready',
             'stores issued', 'stores read', 'exit']
    print(name, 'kernel span %.1f us; CTA start times (us) p0/p50/p100: %.1f %.1f %.1f' % (
        (t[:, 7].max() - t0) / 1e3, 0, np.median(t[:, 0] - t0) / 1e3, (t[:, 0].max() - t0) / 1e3))
    for i in range(1, 8):
        print('   %-24s median +%.2f us  p90 +%.2f us' % (names[i], np.median(rel[:, i]) / 1e3,
                                                          np.percentile(rel[:, i], 90) / 1e3))
