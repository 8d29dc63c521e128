"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, and
exports every symbol include/taiyaki_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from taiyaki_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def libpath():
    return _lib.build()


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'taiyaki_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b([a-z_0-9]+)\s*\(', src)
    return sorted(set(n for n in names if n.startswith('ty_') or n.endswith(('_grad', '_cost'))))


def test_header_lists_expected_entry_points():
    syms = header_symbols()
    for name in ['crf_flipflop_grad', 'crf_flipflop_cost', 'cat_mod_flipflop_grad',
                 'cat_mod_flipflop_cost', 'ty_crf_flipflop', 'ty_flipflop_logz',
                 'ty_flipflop_indices', 'ty_lstm_forward', 'ty_lstm_backward',
                 'ty_gru_forward', 'ty_gru_backward', 'ty_rnn_forward_ex',
                 'ty_rnn_backward_ex', 'ty_col2im_time_major', 'ty_flipflop_train_loss']:
        assert name in syms


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_binding_covers_header(libpath):
    assert sorted(_lib.EXPORTS) == header_symbols()
    assert _lib.lib().ty_version().decode().startswith('taiyaki_b200')


def test_workspace_queries_need_no_gpu(libpath):
    lib = _lib.lib()
    n = lib.ty_crf_flipflop_workspace_bytes(40, 800, 64, 440, 1)
    assert n >= 2 * 64 * 800 * 440 * 4
    assert lib.ty_crf_flipflop_workspace_bytes(40, 800, 64, 440, 0) < 4096
    assert lib.ty_flipflop_logz_workspace_bytes(4, 800, 64) >= 2 * 800 * 64 * 8 * 4
    assert lib.ty_rnn_reserve_bytes(0, 10, 4, 256) == 10 * 4 * 5 * 256 * 4
    assert lib.ty_rnn_reserve_bytes(1, 10, 4, 256) == 10 * 4 * 5 * 256 * 4


def test_ops_refuse_cpu_tensors(libpath):
    import torch
    from taiyaki_b200 import ctc, layers
    x = torch.zeros(4, 1, 40)
    with pytest.raises(_lib.TaiyakiB200Error):
        ctc.crf_flipflop_loss(x, torch.tensor([0, 1]), torch.tensor([2]), 1.0)
    with pytest.raises(_lib.TaiyakiB200Error):
        layers.flipflop_logpartition(x)
    with pytest.raises(_lib.TaiyakiB200Error):
        layers.Lstm(8, 64)(torch.zeros(3, 2, 8))


def test_sass_is_sm100a_only(libpath):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs
