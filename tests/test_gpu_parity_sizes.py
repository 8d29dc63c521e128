"""GPU parity at the BASELINE sizes and at the kernel's internal boundaries.

Round-1 review: every cat-mod case had L <= ~91 (one DP warp at P=8), config A
compared one chunk of 64, P=16 and the unstaged posterior were never checked.
Everything here compares with the REFERENCE'S OWN C (`oracle/_ref`,
c_crf_flipflop.c:434-516, c_cat_mod_flipflop.c:493-582; falls back to the
restatement only when `_ref` was not shipped) and with the fp64 restatement,
at north_star's 1e-4 relative.  Absolute floors: a gradient row is a posterior
divided by nblk, so x/nblk is x of a row's unit mass.

Floors.  fp32 round-off of the recursion grows with the chain length, for the
reference's C exactly as for the kernels (profiles/r2_parity_table.md, measured
on B200 by tools/parity_table.py; distances from fp64 in units of a row's mass):

    nblk    reference C vs fp64      kernels vs fp64       kernels vs reference C
     800    1.2e-5                   1.1e-5                1.4e-5
    1300    1.8e-5                   3.0e-5                2.9e-5
    2000    1.2e-4 .. 2.3e-4         1.1e-4 .. 1.6e-4      1.6e-4 .. 3.6e-4
    2300    3.7e-4                   2.8e-4                5.7e-4
    5000    5.2e-4                   6.2e-4                5.9e-4

so "within 1e-4 relative of the reference" is checked above a floor of 5e-6 of a
row's mass up to 1000 blocks, 5e-5 up to 1500 and 1e-3 beyond (twice that when
both sides are fp32), and `assert_not_noisier` additionally bounds the kernels'
distance from fp64 by a multiple of the reference's own distance (at configs A
and B the kernels are the closer of the two).
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4
OFF = np.array([0, 1, 3, 4, 5], dtype=np.int32)          # A, C(+5mC), G, T
WEIGHTS = np.array([1.0, 1.0, 0.6, 1.0, 1.0], dtype=np.float32)


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


def _impl(oracle):
    return 'ref' if oracle.have_ref() else 'f32'


def _mod_cats(raw, seed=0):
    """Every C (label 1) is 5mC with probability one half."""
    rng = np.random.RandomState(seed)
    return np.concatenate([(r == 1).astype(np.int64) * rng.randint(0, 2, size=len(r))
                           for r in raw])


def assert_not_noisier(ours, ref32, f64, nblk, factor=2.5):
    """max and rms distance from fp64 no more than `factor` times the reference C's."""
    do, dr = np.abs(ours - f64) * nblk, np.abs(ref32 - f64) * nblk
    assert do.max() <= factor * dr.max() + 5e-6, (do.max(), dr.max())
    rms_o, rms_r = np.sqrt((do ** 2).mean()), np.sqrt((dr ** 2).mean())
    assert rms_o <= factor * rms_r + 1e-7, (rms_o, rms_r)


def floor_for(nblk):
    """Absolute floor as a fraction of a row's mass (see the module docstring)."""
    return (5e-6 if nblk <= 1000 else 5e-5 if nblk <= 1500 else 1e-3) / nblk


def _gpu_cat_mod(dev, scores, seqs, seqlen, mod_cats, sharp):
    from taiyaki_b200 import ctc
    x = torch.tensor(scores, device=dev, requires_grad=True)
    cost = ctc.cat_mod_flipflop_loss(x, torch.tensor(seqs), torch.tensor(seqlen),
                                     torch.tensor(mod_cats), OFF, WEIGHTS, sharp)
    cost.sum().backward()
    torch.cuda.synchronize()
    return cost.detach().cpu().numpy(), x.grad.cpu().numpy()


def _gpu_crf(dev, scores, seqs, seqlen, sharp=1.0):
    from taiyaki_b200 import ctc
    cost, grad = ctc.crf_flipflop_cost_grad(torch.tensor(scores, device=dev), torch.tensor(seqs),
                                            torch.tensor(seqlen), sharp, True)
    torch.cuda.synchronize()
    return cost.cpu().numpy(), grad.cpu().numpy()


# --------------------------------------------------------------------------
# (i) cat-mod beyond one DP warp, config B
# --------------------------------------------------------------------------
@pytest.mark.parametrize('nblk,lengths', [
    (700, [300, 257, 256, 255]),            # 2 DP warps at P=8; warp boundary at 256
    (1300, [600, 513, 512, 511, 1, 40]),    # 3 warps, ragged, boundary at 512
    (2300, [2048, 300]),                    # last length of P=8
    (2300, [2049, 2047]),                   # first length of P=16
])
def test_cat_mod_multiwarp_vs_reference(dev, oracle, nblk, lengths):
    nbatch = len(lengths)
    scores = oracle.synth_scores(nblk, nbatch, 45, seed=nblk + nbatch)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, seed=nblk, lengths=lengths)
    mc = _mod_cats(raw, seed=nblk)
    sharp = 1.3
    c_ref, g_ref = oracle.cat_mod_flipflop_loss(scores, seqs, seqlen, mc, OFF, WEIGHTS, sharp,
                                                impl=_impl(oracle))
    c64, g64 = oracle.cat_mod_flipflop_loss(scores, seqs, seqlen, mc, OFF, WEIGHTS, sharp,
                                            impl='f64')
    cost, grad = _gpu_cat_mod(dev, scores, seqs, seqlen, mc, sharp)
    np.testing.assert_allclose(cost, c_ref, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(cost, c64, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g64, rtol=RTOL, atol=floor_for(nblk))
    # fp32 vs fp32: both sides carry round-off, hence twice the floor
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=2 * floor_for(nblk))
    assert_not_noisier(grad, g_ref, g64, nblk)


def test_cat_mod_config_b_full(dev, oracle):
    """BASELINE configs[2]: mGru_cat_mod_flipflop, nblk 2000, 64 chunks, S 45,
    L ~ 440 -- ALL chunks against the reference C, separate operator and the
    fused training loss."""
    from taiyaki_b200 import ctc
    nblk, nbatch = 2000, 64
    scores = oracle.synth_scores(nblk, nbatch, 45, seed=4)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, stride=2, seed=5)
    assert 380 < seqlen.mean() < 500
    mc = _mod_cats(raw, seed=6)
    c_ref, g_ref = oracle.cat_mod_flipflop_loss(scores, seqs, seqlen, mc, OFF, WEIGHTS, 1.0,
                                                impl=_impl(oracle))
    c64, g64 = oracle.cat_mod_flipflop_loss(scores, seqs, seqlen, mc, OFF, WEIGHTS, 1.0,
                                            impl='f64')
    cost, grad = _gpu_cat_mod(dev, scores, seqs, seqlen, mc, 1.0)
    np.testing.assert_allclose(cost, c_ref, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g64, rtol=RTOL, atol=floor_for(nblk))
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=2 * floor_for(nblk))
    assert_not_noisier(grad, g_ref, g64, nblk)
    # the reference's invariants at this size: canonical columns of a row are a posterior
    np.testing.assert_allclose(-grad[:, :, :40].sum(-1) * nblk, 1.0, atol=3e-5)
    # fused loss = cat-mod cost + logZ/nblk (train_flipflop.py:163-182)
    x = torch.tensor(scores, device=dev, requires_grad=True)
    loss = ctc.flipflop_train_loss(x, torch.tensor(seqs), torch.tensor(seqlen), 1.0,
                                   mod_cats=torch.tensor(mc), can_mods_offsets=OFF,
                                   mod_cat_weights=WEIGHTS)
    loss.sum().backward()
    lz, gz = oracle.c_flipflop_logz(np.ascontiguousarray(scores[:, :, :40]), impl='f64')
    np.testing.assert_allclose(loss.detach().cpu().numpy(), c_ref + lz / nblk, rtol=RTOL,
                               atol=1e-5)
    # the fused gradient is a DIFFERENCE of two posteriors (-G + P)/nblk: each within 1e-4
    want = g64.copy()
    want[:, :, :40] += gz / nblk
    tol = RTOL * np.abs(g64)
    tol[:, :, :40] += RTOL * np.abs(gz) / nblk
    err = np.abs(x.grad.cpu().numpy() - want)
    assert np.all(err <= tol + 2 * floor_for(nblk)), (err - tol).max() * nblk


# --------------------------------------------------------------------------
# (ii) config A, every chunk
# --------------------------------------------------------------------------
def test_config_a_all_chunks_vs_reference(dev, oracle):
    from taiyaki_b200 import ctc
    nblk, nbatch = 800, 64
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=0)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, stride=5, seed=1)
    c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl=_impl(oracle))
    cost, grad = _gpu_crf(dev, scores, seqs, seqlen)
    np.testing.assert_allclose(cost, c_ref, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=5e-6 / nblk)
    x = torch.tensor(scores, device=dev, requires_grad=True)
    loss = ctc.flipflop_train_loss(x, torch.tensor(seqs), torch.tensor(seqlen), 1.0)
    loss.sum().backward()
    lz, gz = oracle.c_flipflop_logz(scores, impl='f64')
    np.testing.assert_allclose(loss.detach().cpu().numpy(), c_ref + lz / nblk, rtol=RTOL,
                               atol=1e-5)
    # fused gradient = (-G + P)/nblk, a difference of two posteriors: each term within 1e-4
    # of its own magnitude (relative to the difference the bar would be ill-posed)
    err = np.abs(x.grad.cpu().numpy() - (g_ref + gz / nblk))
    tol = RTOL * (np.abs(g_ref) + np.abs(gz) / nblk) + 2 * floor_for(nblk)
    assert np.all(err <= tol), (err - tol).max() * nblk


def test_reference_speed_test_generator(dev, oracle):
    """The reference's own SPEED_TEST inputs (c_crf_flipflop.c:802-833):
    scores ~ U(-5,5), L_b = nblk*(1+(b-N/2)/(5N))/2."""
    nblk, nbatch = 800, 64
    rng = np.random.RandomState(11)
    scores = rng.uniform(-5, 5, size=(nblk, nbatch, 40)).astype(np.float32)
    lengths = [int(nblk * (1.0 + (b - nbatch / 2) / (5.0 * nbatch)) / 2) for b in range(nbatch)]
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=12, lengths=lengths)
    c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl=_impl(oracle))
    cost, grad = _gpu_crf(dev, scores, seqs, seqlen)
    np.testing.assert_allclose(cost, c_ref, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=5e-6 / nblk)


# --------------------------------------------------------------------------
# (iii) P = 16 and the posterior without shared-memory staging
# --------------------------------------------------------------------------
@pytest.mark.parametrize('nblk,lengths', [
    (4500, [3969, 400]),          # first length above P=8's last warp count at 31 warps
    (5000, [4400, 4000]),         # config-E kernel sweep point (nblk 8000 has L ~ 4400)
    (5400, [4879, 4100]),
])
def test_p16_unstaged_posterior_vs_fp64(dev, oracle, nblk, lengths):
    """Chains of 4000-4900 positions: fp32 round-off of the recursion itself is
    visible at this length (the reference's fp32 C is ~1e-3 relative from fp64 on
    small posterior entries), so the bar is the one of
    test_very_long_chunk_against_fp64: cost at 1e-4, gradient at 1e-4 with a
    floor of 2e-3 of a row's mass, and never more than 3x further from fp64 than
    the reference C."""
    nbatch = len(lengths)
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=nblk)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=nblk + 1, lengths=lengths)
    c64, g64 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='f64')
    c32, g32 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl=_impl(oracle))
    cost, grad = _gpu_crf(dev, scores, seqs, seqlen)
    np.testing.assert_allclose(cost, c64, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g64, rtol=RTOL, atol=2e-3 / nblk)
    np.testing.assert_allclose(grad.sum(2), -1.0 / nblk, rtol=1e-5)
    ours, theirs = np.abs(grad - g64).max(), np.abs(g32 - g64).max()
    assert ours <= 3 * theirs + 1e-9, (ours, theirs)


# --------------------------------------------------------------------------
# (iv) cat-mod through the host-pointer C ABI (libctc.pxd:14-25)
# --------------------------------------------------------------------------
def test_host_pointer_cat_mod_abi(dev, oracle):
    from taiyaki_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    FP, ZP, IP = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_size_t),
                  ctypes.POINTER(ctypes.c_int32))
    SZ = ctypes.c_size_t
    nblk, lengths = 500, [300, 25, 1, 140, 233]
    nbatch = len(lengths)
    scores = oracle.synth_scores(nblk, nbatch, 45, seed=18)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, seed=19, lengths=lengths)
    mc = _mod_cats(raw, seed=20)
    mv, st = oracle.build_indices(seqs, seqlen, 4)
    mm, mf = oracle.build_mod_indices(seqs, seqlen, mc, OFF, WEIGHTS, 4)
    mv, st, mm = (np.ascontiguousarray(a, dtype=np.uintp) for a in (mv, st, mm))
    mf = np.ascontiguousarray(mf, dtype=np.float32)
    sl = seqlen.astype(np.int32)
    score = np.zeros(nbatch, np.float32)
    grad = np.zeros_like(scores)
    lib.cat_mod_flipflop_grad.restype = None
    lib.cat_mod_flipflop_grad.argtypes = [FP, SZ, SZ, SZ, ZP, ZP, ZP, FP, IP, FP, FP]
    lib.cat_mod_flipflop_grad(scores.ctypes.data_as(FP), 45, nblk, nbatch,
                              mv.ctypes.data_as(ZP), st.ctypes.data_as(ZP),
                              mm.ctypes.data_as(ZP), mf.ctypes.data_as(FP),
                              sl.ctypes.data_as(IP), score.ctypes.data_as(FP),
                              grad.ctypes.data_as(FP))
    s_ref, g_ref = oracle.c_cat_mod_flipflop_grad(scores, mv, st, mm, mf, seqlen, _impl(oracle))
    np.testing.assert_allclose(score, s_ref, rtol=RTOL)
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=5e-6)
    score2 = np.zeros(nbatch, np.float32)
    lib.cat_mod_flipflop_cost.restype = None
    lib.cat_mod_flipflop_cost.argtypes = [FP, SZ, SZ, SZ, ZP, ZP, ZP, FP, IP, FP]
    lib.cat_mod_flipflop_cost(scores.ctypes.data_as(FP), 45, nblk, nbatch,
                              mv.ctypes.data_as(ZP), st.ctypes.data_as(ZP),
                              mm.ctypes.data_as(ZP), mf.ctypes.data_as(FP),
                              sl.ctypes.data_as(IP), score2.ctypes.data_as(FP))
    np.testing.assert_allclose(
        score2, oracle.c_cat_mod_flipflop_cost(scores, mv, st, mm, mf, seqlen, _impl(oracle)),
        rtol=RTOL)


# --------------------------------------------------------------------------
# (v) partition function gradient at north_star's tolerance, against fp64
# --------------------------------------------------------------------------
@pytest.mark.parametrize('nblk,nbatch', [(120, 4), (800, 64), (2000, 8)])
def test_logz_gradient_1e4_vs_fp64(dev, oracle, nblk, nbatch):
    """d logZ / d scores are posterior transition probabilities (rows sum to
    one).  1e-4 relative with a floor of 3e-6 of a row's mass -- the floor the
    round-1 test already used; the relative bar is back at north_star's."""
    from taiyaki_b200 import layers
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=nblk)
    x = torch.tensor(scores, device=dev, requires_grad=True)
    lz = layers.flipflop_logpartition(x)
    lz.sum().backward()
    lz64, g64 = oracle.c_flipflop_logz(scores, want_grad=True, impl='f64')
    np.testing.assert_allclose(lz.detach().cpu().numpy(), lz64, rtol=2e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g64, rtol=RTOL, atol=3e-6)
