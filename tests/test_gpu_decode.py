"""Inference lattice operators (taiyaki_b200/decode.py, csrc/viterbi.cu) against golden
vectors generated from the reference's own PyTorch implementations
(taiyaki/decode.py:_flipflop_viterbi, flipflop_make_trans; tests/golden/make_golden.py
decode) and size-independent properties at a basecalling-sized input."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'decode.npz')


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_viterbi_golden(dev, tag):
    from taiyaki_b200 import decode
    g = np.load(GOLDEN)
    scores = torch.tensor(g[tag + '_scores'], device=dev)
    fwd, tb, path = decode.flipflop_viterbi(scores)
    torch.cuda.synchronize()
    assert fwd.dtype == torch.float32 and tb.dtype == torch.int64 and path.dtype == torch.int64
    np.testing.assert_array_equal(path.cpu().numpy(), g[tag + '_path'])       # bit-exact path
    np.testing.assert_array_equal(tb.cpu().numpy(), g[tag + '_tb'])
    # the reference adds -1e30 + score for the flop states of the first blocks in fp32
    np.testing.assert_allclose(fwd.cpu().numpy(), g[tag + '_fwd'], rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_make_trans_golden(dev, tag):
    from taiyaki_b200 import decode
    g = np.load(GOLDEN)
    scores = torch.tensor(g[tag + '_scores'], device=dev)
    trans = decode.flipflop_make_trans(scores)
    torch.cuda.synchronize()
    np.testing.assert_allclose(trans.cpu().numpy(), g[tag + '_trans'], rtol=1e-4, atol=1e-6)


def test_viterbi_properties_full_size(dev):
    """4000 blocks x 64 chunks: the path is a valid flip-flop path, its score equals the
    best final forward score, and no sampled alternative path beats it."""
    from taiyaki_b200 import decode
    torch.manual_seed(0)
    T, N, nb = 4000, 64, 4
    scores = 3 * torch.randn(T, N, 40, device=dev)
    fwd, tb, path = decode.flipflop_viterbi(scores)
    trans = decode.flipflop_make_trans(scores)
    torch.cuda.synchronize()
    frm, to = path[:-1], path[1:]
    # allowed: any -> flip (to < nb); flip b -> flop b; flop b -> flop b
    ok = (to < nb) | ((to >= nb) & ((frm == to - nb) | (frm == to)))
    assert bool(ok.all())
    idx = torch.where(to < nb, to * 2 * nb + frm, 2 * nb * nb + frm)
    pscore = scores.gather(2, idx.unsqueeze(2)).squeeze(2).sum(0)
    best = fwd[-1].max(1).values
    assert torch.allclose(pscore, best, rtol=1e-5, atol=1e-2)
    assert bool((path[0] < nb).all())                      # flop states start at -1e30
    # posteriors: every block's transition probabilities sum to one
    assert torch.allclose(trans.sum(2), torch.ones(T, N, device=dev), atol=1e-4)
    assert bool((trans >= 0).all())
