"""ctypes binding of libtaiyaki_b200.so (include/taiyaki_b200.h).

The library is built in-tree by `build()` (nvcc, sm_100a only) and loaded
lazily.  There is deliberately no fallback: if the shared object is missing or
a CUDA call fails the operators raise.
"""
import ctypes
import glob
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.environ.get('TY_B200_LIB', os.path.join(_HERE, 'libtaiyaki_b200.so'))   # override: A/B builds

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '-Xcompiler', '-fPIC', '-shared']

c_void_p, c_int, c_float, c_size_t, c_int64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_int64)

_SIGNATURES = {
    'ty_last_error_string': (ctypes.c_char_p, []),
    'ty_version': (ctypes.c_char_p, []),
    'ty_crf_tuning': (None, [c_int, c_int]),
    'ty_crf_last_path': (c_int, []),
    'ty_crf_flipflop_workspace_bytes': (c_size_t, [c_int] * 5),
    'ty_crf_flipflop': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_float, c_int, c_float, c_void_p,
                                c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ty_flipflop_indices': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    'ty_flipflop_indices_checked': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p]),
    'ty_flipflop_logz_workspace_bytes': (c_size_t, [c_int] * 3),
    'ty_flipflop_logz': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                 c_float, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'ty_flipflop_logz_phase': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                       c_float, c_void_p, c_int, c_int, c_void_p, c_size_t, c_int,
                                       c_void_p]),
    'ty_flipflop_train_loss_workspace_bytes': (c_size_t, [c_int] * 5),
    'ty_flipflop_train_loss': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_float, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ty_col2im_time_major': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
    'ty_col2im_time_major_ld': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_void_p, c_void_p]),
    'ty_im2col_time_major_bf16': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_int, c_int, c_void_p, c_void_p]),
    'ty_conv_small_supported': (c_int, [c_int] * 3),
    'ty_conv_in1_supported': (c_int, [c_int] * 3),
    'ty_conv_in1_forward': (c_int, [c_void_p] * 3 + [c_int] * 7 + [c_void_p, c_void_p]),
    'ty_conv_in1_wgrad': (c_int, [c_void_p] * 2 + [c_int] * 7 + [c_void_p] * 3),
    'ty_conv_small_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'ty_conv_small_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    'ty_flipflop_viterbi': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    'ty_flipflop_remap': (c_int, [c_void_p] * 6 + [c_int] * 3 + [ctypes.c_double] + [c_void_p] * 5),
    'ty_batch_counts_len': (c_int, []),
    'ty_sample_chunks': (c_int, [c_void_p] * 9 + [c_int] * 3 + [c_void_p] + [c_int] * 3 +
                         [c_void_p] * 10),
    'ty_rnn_reserve_bytes': (c_size_t, [c_int] * 4),
    'ty_rnn_forward_ex': (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    'ty_rnn_backward_ex': (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                   c_void_p]),
    'ty_rnn_um_supported': (c_int, [c_int]),
    'ty_rnn_forward_um': (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    'ty_rnn_backward_um': (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ty_lstm_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p]),
    'ty_lstm_backward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    'ty_gru_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p]),
    'ty_gemm_bf16': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             c_void_p, c_int, c_int, c_void_p, c_float, c_int, c_int, c_int,
                             c_void_p]),
    'ty_gru_backward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}
# host-pointer drop-ins carrying the reference's own names (libctc.pxd:3-25)
HOST_ABI = ['crf_flipflop_grad', 'crf_flipflop_cost', 'cat_mod_flipflop_grad',
            'cat_mod_flipflop_cost']
EXPORTS = sorted(_SIGNATURES) + HOST_ABI


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, '*.cu')))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into taiyaki_b200/libtaiyaki_b200.so: one object
    per source (in parallel, rebuilt only when the source or a header changed),
    then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = sources()
    headers = glob.glob(os.path.join(_CSRC, '*.cuh')) + [
        os.path.join(os.path.dirname(_HERE), 'include', 'taiyaki_b200.h')]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(_HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != '-shared']
    jobs, objs = [], []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or \
                os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append([nvcc] + compile_flags + ['-c', '-o', obj, src])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            list(pool.map(run, jobs))
    if jobs or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        run([nvcc] + NVCC_FLAGS[:2] + ['-shared', '-o', LIB_PATH] + objs)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'taiyaki_b200: %s is missing -- run `python -c "import __graft_entry__ as g; '
                'g.build()"`; there is no CPU fallback' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class TaiyakiB200Error(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().ty_last_error_string().decode()
        raise TaiyakiB200Error('%s failed (code %d): %s' % (what, rc, msg))


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise TaiyakiB200Error(
            'taiyaki_b200: %s must be a CUDA tensor (got %s); this package has no CPU path'
            % (name, t.device))


#: number of kernels of THIS library launched since import (bench.py reports it)
LAUNCHES = 0
#: when not None, a dict name -> list of (start, end) CUDA events recorded around
#: the named launches on the launching stream (bench.py's roofline leg)
PROFILE = None


def count_launches(n):
    global LAUNCHES
    LAUNCHES += n


class timed:
    """Record CUDA events around a group of launches when PROFILE is enabled."""

    def __init__(self, name, device):
        self.name, self.device = name, device

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.device))

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record(torch.cuda.current_stream(self.device))
            PROFILE.setdefault(self.name, []).append((self.a, self.b))


_workspaces = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per (device, stream); kernels are stream ordered."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf
