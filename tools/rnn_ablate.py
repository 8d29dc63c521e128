#!/usr/bin/env python
"""Ablation timing of the LSTM recurrence kernels: build patched copies of the
library (one experiment each: a part of the step removed) under
build_variants/, then time them on the GPU.  Tells which part of a step costs
what.  Tuning tool, not part of the product.
    python tools/rnn_ablate.py build     (here)
    python tools/rnn_ablate.py run       (on the GPU box)"""
import ctypes
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BV = os.path.join(ROOT, 'build_variants')

# name -> list of (old, new) textual patches applied to csrc/rnn.cu
PATCHES = {
    'base': [],
    'bwd_nostore': [
        ('                    if (a.dxproj16) a.dxproj16[xrow + (size_t)g * H] = __float2bfloat16(dg[g]);\n'
         '                    else __stcs(a.dxproj + xrow + (size_t)g * H, dg[g]);\n',
         '                    if (dg[g] == 123.456f) __stcs(a.dxproj + xrow + (size_t)g * H, dg[g]);\n')],
    'bwd_noload': [
        ('    auto issue_in = [&](int s, int slot) {\n        if (s < T) {',
         '    auto issue_in = [&](int s, int slot) {\n        if (s < kXLook) {')],
    'bwd_halfmma': [
        ('            for (int kp = 0; kp < KT / 2; kp++) {\n                uint32_t bf[4];\n'
         '                ldmatrix_x4(bf, ds_base + ld_off + kp * 64);',
         '            for (int kp = 0; kp < KT / 4; kp++) {\n                uint32_t bf[4];\n'
         '                ldmatrix_x4(bf, ds_base + ld_off + kp * 64);')],
    'bwd_nosync': [
        ('        if (s + 1 < T) {\n            __syncthreads();', '        if (s + 1 < T) {\n            __syncwarp();')],
    'fwd_nostore': [
        ('                if (valid) __stcs(cstate_out + cell, c);', ''),
        ('            if (valid) {\n                __stcs(reinterpret_cast<float4 *>(gates_out) + cell, sv);',
         '            if (valid && hnew[col] == 123.456f) {\n                __stcs(reinterpret_cast<float4 *>(gates_out) + cell, sv);')],
    'fwd_noload': [
        ('    auto issue_x = [&](int s, int slot) {\n        if (s < T) {',
         '    auto issue_x = [&](int s, int slot) {\n        if (s < kXLook) {')],
}


def build():
    os.makedirs(BV, exist_ok=True)
    procs = []
    for name, patches in PATCHES.items():
        src = os.path.join(BV, 'src_' + name)
        shutil.rmtree(src, ignore_errors=True)
        shutil.copytree(os.path.join(ROOT, 'taiyaki_b200', 'csrc'), src)
        path = os.path.join(src, 'rnn.cu')
        text = open(path).read()
        for old, new in patches:
            assert old in text, (name, old[:60])
            text = text.replace(old, new)
        open(path, 'w').write(text)
        out = os.path.join(BV, 'lib_%s.so' % name)
        cmd = ['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
               '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-w', '-I', os.path.join(ROOT, 'include'),
               '-o', out] + sorted(glob.glob(os.path.join(src, '*.cu')))
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        assert p.wait() == 0


def run():
    import torch
    dev = torch.device('cuda:0')
    T, N, H = 800, 64, 256
    torch.manual_seed(0)
    xproj = torch.randn(T, N, 4 * H, device=dev)
    w_hh = torch.randn(4 * H, H, device=dev) / 16
    y = torch.empty(T, N, H, device=dev)
    dy = torch.randn(T, N, H, device=dev)
    dxp = torch.empty(T, N, 4 * H, device=dev)
    reserve = torch.empty(T * N * 5 * H, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    a = torch.randn(4096, 4096, device=dev)
    for _ in range(20):
        (a @ a).sum().item()
    vp = ctypes.c_void_p
    P = lambda t: vp(t.data_ptr())
    for name in PATCHES:
        lib = ctypes.CDLL(os.path.join(BV, 'lib_%s.so' % name))
        fwd, bwd = lib.ty_lstm_forward, lib.ty_lstm_backward
        fwd.restype = bwd.restype = ctypes.c_int
        fwd.argtypes = [vp, vp, vp] + [ctypes.c_int] * 4 + [vp, vp, vp]
        bwd.argtypes = [vp, vp] + [ctypes.c_int] * 4 + [vp, vp, vp, vp, vp]
        st = vp(torch.cuda.current_stream().cuda_stream)
        f = lambda: fwd(P(xproj), None, P(w_hh), T, N, H, 0, P(y), P(reserve), st)
        b = lambda: bwd(P(dy), P(w_hh), T, N, H, 0, P(y), P(reserve), P(dxp), None, st)
        res = []
        for fn in (f, b):
            ts = []
            for it in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn()
                e1.record()
                torch.cuda.synchronize()
                assert rc == 0
                ts.append(e0.elapsed_time(e1))
            ts = sorted(ts[2:])
            res.append(ts[len(ts) // 2])
        print('%-14s fwd %.4f ms (%4.0f cyc/step)   bwd %.4f ms (%4.0f cyc/step)' % (
            name, res[0], res[0] * 1e-3 / T * 1965e6, res[1], res[1] * 1e-3 / T * 1965e6))


if __name__ == '__main__':
    build() if sys.argv[1:] == ['build'] else run()
