#!/usr/bin/env python
"""Time the LSTM forward recurrence of every library variant under
build_variants/ (built with `tools/rnn_ablate.py build`): ablation experiments
that tell which part of the step costs what.  Not part of the product."""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    'base': [],
    'noxload': ['-DTY_ABL_NOXLOAD'],
    'noprefetch': ['-DTY_ABL_NOPREFETCH'],
    'nostore': ['-DTY_ABL_NOSTORE'],
    'halfmma': ['-DTY_ABL_HALFMMA'],
    'nogates': ['-DTY_ABL_NOGATES'],
    'nox_nostore': ['-DTY_ABL_NOXLOAD', '-DTY_ABL_NOSTORE', '-DTY_ABL_NOPREFETCH'],
    'all': ['-DTY_ABL_NOXLOAD', '-DTY_ABL_NOSTORE', '-DTY_ABL_NOPREFETCH', '-DTY_ABL_NOGATES', '-DTY_ABL_HALFMMA'],
}


def build():
    os.makedirs(os.path.join(ROOT, 'build_variants'), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(ROOT, 'taiyaki_b200', 'csrc', '*.cu')))
    procs = []
    for name, flags in VARIANTS.items():
        out = os.path.join(ROOT, 'build_variants', 'lib_%s.so' % name)
        cmd = ['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
               '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-o', out] + flags + srcs
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
    reserve = torch.empty(T * N * 5 * H, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    a = torch.randn(4096, 4096, device=dev)
    for _ in range(20):
        (a @ a).sum().item()
    vp = ctypes.c_void_p
    for name in VARIANTS:
        lib = ctypes.CDLL(os.path.join(ROOT, 'build_variants', 'lib_%s.so' % name))
        fn = lib.ty_lstm_forward
        fn.restype = ctypes.c_int
        fn.argtypes = [vp, vp, vp] + [ctypes.c_int] * 4 + [vp, vp, vp]
        st = vp(torch.cuda.current_stream().cuda_stream)
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(vp(xproj.data_ptr()), None, vp(w_hh.data_ptr()), T, N, H, 0, vp(y.data_ptr()),
                    vp(reserve.data_ptr()), st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        print('%-14s %.4f ms  (%.0f cycles/step at 1965 MHz)' % (name, ts[len(ts) // 2], ts[len(ts) // 2] * 1e-3 / T * 1965e6))


if __name__ == '__main__':
    build() if sys.argv[1:] == ['build'] else run()
