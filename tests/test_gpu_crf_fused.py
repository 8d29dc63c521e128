"""GPU tests of the fused chain + posterior kernel (csrc/crf_fused.cu) against the chain
kernel + posterior kernel pair it replaces (csrc/crf_flipflop.cu) and against the fp64
restatement of the reference (oracle/, c_crf_flipflop.c:434-516, c_cat_mod_flipflop.c:493-582):
the shapes where the two chains meet (one, two, three blocks; odd block counts), chunks
without labels, chunk lengths at the thread-block boundaries, the cat-mod variant, and the
selection between the two paths.  `ty_crf_last_path` says which kernels a call launched, so
a silent fall-back to the other path fails the test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OFF = np.array([0, 1, 3, 4, 5], dtype=np.int32)          # A, C(+5mC), G, T
WEIGHTS = np.array([1.0, 1.0, 0.6, 1.0, 1.0], dtype=np.float32)
FUSED, TWO_KERNEL = 2, 1


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


@pytest.fixture(autouse=True)
def default_tuning():
    from taiyaki_b200 import _lib
    yield
    _lib.lib().ty_crf_tuning(0, 1)


def _inputs(nblk, nbatch, mod, seed, lengths=None):
    from oracle import oracle
    S = 45 if mod else 40
    scores = oracle.synth_scores(nblk, nbatch, S, seed=seed)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, stride=5, seed=seed + 1, lengths=lengths)
    mod_cats = None
    if mod:
        rng = np.random.RandomState(seed + 2)
        mod_cats = np.concatenate([(r == 1).astype(np.int64) * rng.randint(0, 2, size=len(r)) for r in raw]
                                  + [np.zeros(0, np.int64)])
    return scores, seqs, seqlen, mod_cats


def _run(dev, scores, seqs, seqlen, mod_cats, sharp, fused):
    from taiyaki_b200 import _lib, ctc
    lib = _lib.lib()
    lib.ty_crf_tuning(0, int(fused))
    x = torch.tensor(scores, device=dev, requires_grad=True)
    if mod_cats is None:
        cost = ctc.crf_flipflop_loss(x, torch.tensor(seqs), torch.tensor(seqlen), sharp)
    else:
        cost = ctc.cat_mod_flipflop_loss(x, torch.tensor(seqs), torch.tensor(seqlen), torch.tensor(mod_cats),
                                         OFF, WEIGHTS, sharp)
    path = lib.ty_crf_last_path()
    cost.sum().backward()
    torch.cuda.synchronize()
    return cost.detach().cpu().numpy(), x.grad.cpu().numpy(), path


def _oracle(scores, seqs, seqlen, mod_cats, sharp):
    from oracle import oracle
    if mod_cats is None:
        return oracle.crf_flipflop_loss(scores, seqs, seqlen, sharp, impl='f64')
    return oracle.cat_mod_flipflop_loss(scores, seqs, seqlen, mod_cats, OFF, WEIGHTS, sharp, impl='f64')


@pytest.mark.parametrize('mod', [False, True])
@pytest.mark.parametrize('nblk,nbatch,lengths', [
    (1, 3, [1, 1, 1]), (2, 3, [1, 2, 1]), (3, 4, [2, 1, 3, 2]), (4, 2, [3, 2]), (5, 5, None), (7, 3, None),
    (33, 6, None), (64, 9, None), (65, 4, [1, 36, 20, 65]), (200, 5, None), (401, 3, None)])
def test_fused_matches_two_kernel_path_and_fp64(dev, nblk, nbatch, lengths, mod):
    scores, seqs, seqlen, mod_cats = _inputs(nblk, nbatch, mod, seed=nblk, lengths=lengths)
    c1, g1, p1 = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=True)
    c0, g0, p0 = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=False)
    assert (p1, p0) == (FUSED, TWO_KERNEL)
    np.testing.assert_allclose(c1, c0, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(g1, g0, rtol=2e-5, atol=2e-7 / nblk)
    c64, g64 = _oracle(scores, seqs, seqlen, mod_cats, 1.0)
    np.testing.assert_allclose(c1, c64, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g1, g64, rtol=1e-4, atol=5e-6 / nblk)
    # rows are posteriors: the canonical transitions of a row sum to -1/nblk
    np.testing.assert_allclose(g1[:, :, :40].sum(2), -1.0 / nblk, rtol=1e-5)


@pytest.mark.parametrize('mod', [False, True])
def test_chunks_without_labels_in_the_batch(dev, mod):
    """seqlen 0 (c_crf_flipflop.c:269-272, :458-464): zero cost and gradient for that chunk, the
    others unaffected; both CTAs of the chunk's cluster leave before the first barrier."""
    scores, seqs, seqlen, mod_cats = _inputs(50, 5, mod, seed=3, lengths=[20, 0, 31, 0, 7])
    c1, g1, p1 = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=True)
    c0, g0, p0 = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=False)
    assert (p1, p0) == (FUSED, TWO_KERNEL)
    assert c1[1] == 0 and c1[3] == 0 and not g1[:, 1].any() and not g1[:, 3].any()
    np.testing.assert_allclose(c1, c0, rtol=1e-6)
    np.testing.assert_allclose(g1, g0, rtol=2e-5, atol=4e-9)


@pytest.mark.parametrize('L', [127, 128, 129, 511, 512, 513, 1024, 1100])
def test_chunk_lengths_at_the_block_boundaries(dev, L):
    """positions per thread x warps: a DP warp boundary falls at multiples of 128 (P = 4); the
    fused kernel takes chunks up to about 1100 positions (shared memory of its eight posterior
    warps), the kernel pair the longer ones."""
    nblk = L + 40
    scores, seqs, seqlen, _ = _inputs(nblk, 2, False, seed=L, lengths=[L, max(1, L // 3)])
    c1, g1, p1 = _run(dev, scores, seqs, seqlen, None, 1.0, fused=True)
    c0, g0, p0 = _run(dev, scores, seqs, seqlen, None, 1.0, fused=False)
    assert (p1, p0) == (FUSED, TWO_KERNEL)
    np.testing.assert_allclose(c1, c0, rtol=1e-6)
    # L close to nblk leaves few alignments and amplifies fp32 round-off in EVERY fp32 implementation
    # (tools/crf_fused_diag.py on B200, distance from fp64 in units of a row's mass at L 1024: fused
    # 1.7e-4, kernel pair 1.7e-4, the reference's C 1.3e-4): the fused kernel is held to twice the
    # larger of the kernel pair's and the reference C's distance
    from oracle import oracle
    c64, g64 = _oracle(scores, seqs, seqlen, None, 1.0)
    _, g32 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='ref' if oracle.have_ref() else 'f32')
    np.testing.assert_allclose(c1, c64, rtol=1e-4)
    d1, d0, dr = (np.abs(g - g64).max() * nblk for g in (g1, g0, g32))
    assert d1 <= 2 * max(d0, dr) + 5e-6, (d1, d0, dr)
    r1, r0, rr = (np.sqrt(((g - g64) ** 2).mean()) * nblk for g in (g1, g0, g32))
    assert r1 <= 2 * max(r0, rr) + 1e-7, (r1, r0, rr)


def test_sharpening_and_scaling(dev):
    scores, seqs, seqlen, _ = _inputs(120, 4, False, seed=9)
    for sharp in (0.5, 2.0):
        c1, g1, p1 = _run(dev, scores, seqs, seqlen, None, sharp, fused=True)
        c64, g64 = _oracle(scores, seqs, seqlen, None, sharp)
        assert p1 == FUSED
        np.testing.assert_allclose(c1, c64, rtol=1e-4)
        np.testing.assert_allclose(g1, g64, rtol=1e-4, atol=5e-6 / 120)


@pytest.mark.parametrize('fused', [True, False])
def test_cat_mod_prior_weights_with_zero_entries(dev, fused, monkeypatch):
    """Category weights as --mod_prior_factor produces them (alphabet.py:68-100): 0 for a canonical
    base without modifications, odds around 1 for C and 5mC."""
    import sys
    weights = np.array([0.0, 1.0157232, 0.98452014, 0.0, 0.0], dtype=np.float32)
    monkeypatch.setattr(sys.modules[__name__], 'WEIGHTS', weights)
    scores, seqs, seqlen, mod_cats = _inputs(90, 6, True, seed=17)
    c1, g1, p1 = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=fused)
    assert p1 == (FUSED if fused else TWO_KERNEL)
    c64, g64 = _oracle(scores, seqs, seqlen, mod_cats, 1.0)
    np.testing.assert_allclose(c1, c64, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g1, g64, rtol=1e-4, atol=5e-6 / 90)


def test_long_chunks_take_the_two_kernel_path(dev):
    """Rows that do not fit in shared memory next to the ring (from about 1100 positions on) keep
    the spill + posterior kernel pair."""
    L = 1300
    scores, seqs, seqlen, _ = _inputs(L + 10, 1, False, seed=5, lengths=[L])
    c1, g1, p1 = _run(dev, scores, seqs, seqlen, None, 1.0, fused=True)
    assert p1 == TWO_KERNEL
    assert np.isfinite(c1).all() and np.isfinite(g1).all()


def test_repeated_calls_are_bit_identical(dev):
    """No atomics on the gradient path: the summation order is fixed."""
    scores, seqs, seqlen, mod_cats = _inputs(300, 8, True, seed=21)
    c1, g1, _ = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=True)
    for _ in range(3):
        c2, g2, p = _run(dev, scores, seqs, seqlen, mod_cats, 1.0, fused=True)
        assert p == FUSED
        assert np.array_equal(c1, c2) and np.array_equal(g1, g2)
