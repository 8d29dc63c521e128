"""GPU parity tests for the CRF loss path: the CUDA kernels behind the C ABI
against the CPU oracle (oracle/) and the golden vectors generated from the
reference (tests/golden/).  Tolerance (north_star): 1e-4 relative fp32 on loss
and gradients; absolute floors are stated per test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from taiyaki_b200 import _lib
    _lib.lib()      # fail loudly when the extension is missing
    return torch.device('cuda:0')


def gpu_crf(dev, scores, seqs, seqlen, sharp=1.0, want_grad=True):
    from taiyaki_b200 import ctc
    x = torch.tensor(scores, device=dev)
    cost, grad = ctc.crf_flipflop_cost_grad(x, torch.tensor(seqs), torch.tensor(seqlen),
                                            sharp, want_grad)
    torch.cuda.synchronize()
    return cost.cpu().numpy(), (grad.cpu().numpy() if want_grad else None)


def raw_crf(dev, lp, move, stay, seqlen, modmove=None, modfact=None, want_grad=True):
    """Straight through the device C ABI with the reference's raw conventions."""
    from taiyaki_b200 import _lib
    lib = _lib.lib()
    x = torch.tensor(lp, device=dev)
    nblk, nbatch, ntrans = x.shape
    n = max(int(np.sum(seqlen)), 1)

    def idx(a, dt=torch.int32):
        buf = torch.zeros(n, dtype=dt, device=dev)
        a = np.asarray(a)
        buf[:len(a)] = torch.tensor(a.astype(np.float32 if dt == torch.float32 else np.int64),
                                    device=dev).to(dt)
        return buf
    mv, st = idx(move), idx(stay)
    mm = idx(modmove) if modmove is not None else None
    mf = idx(modfact, torch.float32) if modfact is not None else None
    sl = torch.tensor(np.asarray(seqlen), device=dev, dtype=torch.int32)
    max_len = int(np.max(seqlen))
    score = torch.empty(nbatch, device=dev)
    grad = torch.empty_like(x) if want_grad else None
    ws = _lib.workspace(lib.ty_crf_flipflop_workspace_bytes(ntrans, nblk, nbatch, max_len,
                                                            int(want_grad)), dev)
    rc = lib.ty_crf_flipflop(_lib.ptr(x), ntrans, nblk, nbatch, _lib.ptr(mv), _lib.ptr(st),
                             _lib.ptr(mm), _lib.ptr(mf), _lib.ptr(sl), max_len, 1.0, ntrans,
                             1.0, _lib.ptr(score), 1.0, _lib.ptr(grad), _lib.ptr(ws),
                             ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, 'ty_crf_flipflop')
    torch.cuda.synchronize()
    return score.cpu().numpy(), (grad.cpu().numpy() if want_grad else None)


def test_kat_crf_twostate(dev, kat):
    # c_crf_flipflop.c:520-695: forward = backward = -2.378088
    sc, gr = raw_crf(dev, kat['crf_logprob'], kat['crf_move'], kat['crf_stay'], kat['crf_seqlen'])
    np.testing.assert_allclose(sc, -2.378088, atol=3e-6)
    np.testing.assert_allclose(gr, kat['crf_grad'], rtol=RTOL, atol=2e-6)
    sc2, _ = raw_crf(dev, kat['crf_logprob'], kat['crf_move'], kat['crf_stay'],
                     kat['crf_seqlen'], want_grad=False)
    np.testing.assert_allclose(sc2, -2.378088, atol=3e-6)


def test_kat_cat_mod(dev, kat):
    # c_cat_mod_flipflop.c:586-870: -52.354622 / -195.435257
    sc, gr = raw_crf(dev, kat['cm_logprob'], kat['cm_move'], kat['cm_stay'], kat['cm_seqlen'],
                     kat['cm_modmove'], kat['cm_modfact'])
    np.testing.assert_allclose(sc, [-52.354622, -195.435257], rtol=2e-6)
    np.testing.assert_allclose(gr, kat['cm_grad'], rtol=RTOL, atol=2e-5)


def test_unit_ctc_loss_case(dev, kat):
    # test/unit/test_ctc_loss.py:84-103 through the operator API
    from taiyaki_b200 import ctc, layers
    scores = torch.tensor(kat['unit_scores'], device=dev)
    assert abs(float(layers.log_partition_flipflop(scores))) < 1e-5
    for seq, prob in zip(kat['unit_seqs'], kat['unit_probs']):
        loss = ctc.crf_flipflop_loss(scores, torch.tensor(seq), torch.tensor([3]), 1.0)
        assert abs(float(torch.exp(-loss * 4)) - prob) < 1e-6


def test_unit_ctc_gradient_check(dev, kat):
    # test/unit/test_ctc_loss.py:105-135 (autograd path)
    from taiyaki_b200 import ctc
    torch.manual_seed(0)
    for seq in kat['unit_seqs'][:2]:
        x = torch.tensor(kat['unit_scores'], device=dev, requires_grad=True)
        loss = ctc.crf_flipflop_loss(x, torch.tensor(seq), torch.tensor([3]), 1.0).sum()
        loss.backward()
        dx = torch.randn_like(x) * 1e-3
        loss2 = ctc.crf_flipflop_loss(x.detach() + dx, torch.tensor(seq), torch.tensor([3]),
                                      1.0).sum()
        est = float((dx * x.grad).sum())
        assert abs(float(loss2 - loss) / float(loss) - est / float(loss)) < 1e-5


@pytest.mark.parametrize('tag', ['a', 'b', 'e'])
def test_golden_random_crf(dev, golden_random, tag):
    g = golden_random
    scores, seqs, seqlen = g[tag + '_scores'], g[tag + '_seqs'], g[tag + '_seqlen']
    sharp = float(g[tag + '_sharp'])
    nblk = scores.shape[0]
    cost, grad = gpu_crf(dev, scores, seqs, seqlen, sharp)
    np.testing.assert_allclose(cost, -g[tag + '_score'] / nblk / sharp, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, -g[tag + '_grad'] / nblk, rtol=RTOL, atol=2e-6 / nblk)
    cost2, _ = gpu_crf(dev, scores, seqs, seqlen, sharp, want_grad=False)
    np.testing.assert_allclose(cost2, -g[tag + '_score_costonly'] / nblk / sharp, rtol=RTOL,
                               atol=1e-6)


@pytest.mark.parametrize('tag', ['c', 'd'])
def test_golden_random_cat_mod(dev, golden_random, tag):
    from taiyaki_b200 import ctc
    g = golden_random
    scores, seqs, seqlen = g[tag + '_scores'], g[tag + '_seqs'], g[tag + '_seqlen']
    sharp = float(g[tag + '_sharp'])
    nblk = scores.shape[0]
    x = torch.tensor(scores, device=dev, requires_grad=True)
    cost = ctc.cat_mod_flipflop_loss(x, torch.tensor(seqs), torch.tensor(seqlen),
                                     torch.tensor(g[tag + '_mod_cats']),
                                     g[tag + '_can_mods_offsets'], g[tag + '_mod_cat_weights'],
                                     sharp)
    cost.sum().backward()
    np.testing.assert_allclose(cost.detach().cpu().numpy(), -g[tag + '_score'] / nblk / sharp,
                               rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), -g[tag + '_grad'] / nblk, rtol=RTOL,
                               atol=2e-6 / nblk)


def test_indices_kernel_matches_flipflopfings(dev, oracle):
    from taiyaki_b200 import ctc
    seqs, seqlen, raw = oracle.synth_seqs(300, 9, stride=5, seed=3,
                                          lengths=[40, 1, 0, 33, 2, 150, 0, 7, 64])
    mod_cats = np.concatenate([(r == 1).astype(np.int64) for r in raw])
    off = np.array([0, 1, 3, 4, 5], dtype=np.int32)
    w = np.array([1.0, 1.0, 0.5, 1.0, 1.0], dtype=np.float32)
    mv, st, sl, mm, mf, max_len = ctc.build_indices(seqs, seqlen, 4, dev, mod_cats, off, w)
    torch.cuda.synchronize()
    emv, est = oracle.build_indices(seqs, seqlen, 4)
    emm, emf = oracle.build_mod_indices(seqs, seqlen, mod_cats, off, w, 4)
    assert max_len == 150
    np.testing.assert_array_equal(st.cpu().numpy()[:len(est)], est)
    np.testing.assert_array_equal(mv.cpu().numpy()[:len(emv)], emv)
    np.testing.assert_array_equal(mm.cpu().numpy()[:len(emm)], emm)
    np.testing.assert_array_equal(mf.cpu().numpy()[:len(emf)], emf)
    np.testing.assert_array_equal(sl.cpu().numpy(), seqlen)


@pytest.mark.parametrize('nblk,nbatch,lengths', [
    (120, 6, [40, 1, 52, 33, 64, 47]),
    (300, 5, [160, 131, 1, 150, 2]),          # more than one thread-block tile of warps
    (64, 3, [64, 63, 10]),                    # sequence as long as the chunk
    (1500, 2, [700, 820]),                    # long chain
    (37, 4, [20, 30, 5, 11]),                 # ragged tail of the posterior row tile
])
def test_live_oracle_parity(dev, oracle, nblk, nbatch, lengths):
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=nblk)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=nblk + 1, lengths=lengths)
    impl = 'ref' if oracle.have_ref() else 'f32'
    c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl=impl)
    c64, g64 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='f64')
    cost, grad = gpu_crf(dev, scores, seqs, seqlen, 1.0)
    np.testing.assert_allclose(cost, c_ref, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=5e-6 / nblk)
    # and no further from the fp64 ground truth than 1e-4 either
    np.testing.assert_allclose(grad, g64, rtol=RTOL, atol=5e-6 / nblk)
    np.testing.assert_allclose(cost, c64, rtol=RTOL, atol=1e-6)


def test_very_long_chunk_against_fp64(dev, oracle):
    """3400 blocks x 3100 positions (25 DP warps per chain; the posterior reads
    the alpha / beta rows from global memory because they no longer fit in
    shared memory): fp32 round-off of the recursion itself is visible at this
    length -- the reference's own fp32 C is 1e-3 relative away from fp64 on the
    small posterior entries -- so the check is against fp64 with a floor of 2e-3
    of a row's mass, and the kernel must not be further from fp64 than three
    times the reference C is."""
    nblk, nbatch, lengths = 3400, 2, [3100, 150]
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=nblk)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=nblk + 1, lengths=lengths)
    c64, g64 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='f64')
    impl = 'ref' if oracle.have_ref() else 'f32'
    c32, g32 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl=impl)
    cost, grad = gpu_crf(dev, scores, seqs, seqlen, 1.0)
    np.testing.assert_allclose(cost, c64, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(grad, g64, rtol=RTOL, atol=2e-3 / nblk)
    np.testing.assert_allclose(grad.sum(2), -1.0 / nblk, rtol=1e-5)      # rows are posteriors / nblk
    ours, theirs = np.abs(grad - g64).max(), np.abs(g32 - g64).max()
    assert ours <= 3 * theirs + 1e-9, (ours, theirs)


def test_sharpen_identity_and_scaling(dev, oracle):
    nblk, nbatch = 90, 4
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=5)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=6)
    a, ga = gpu_crf(dev, scores, seqs, seqlen, 2.0)
    b, gb = gpu_crf(dev, 2 * scores, seqs, seqlen, 1.0)
    np.testing.assert_allclose(a, b / 2, rtol=1e-6)
    np.testing.assert_allclose(ga, gb, rtol=1e-5, atol=1e-9)
    c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 2.0, impl='f32')
    np.testing.assert_allclose(a, c_ref, rtol=RTOL)
    np.testing.assert_allclose(ga, g_ref, rtol=RTOL, atol=5e-6 / nblk)


def test_full_size_properties(dev, oracle):
    """BASELINE config A (nblk=800, N=64, S=40): size-independent properties --
    posterior rows sum to one, gradients non-positive, forward-only cost equals
    the forward/backward average, and a sampled chunk matches the oracle."""
    nblk, nbatch = 800, 64
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=0)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, stride=5, seed=1)
    cost, grad = gpu_crf(dev, scores, seqs, seqlen, 1.0)
    cost_f, _ = gpu_crf(dev, scores, seqs, seqlen, 1.0, want_grad=False)
    assert np.all(np.isfinite(cost)) and np.all(np.isfinite(grad))
    np.testing.assert_allclose(-grad.sum(-1) * nblk, 1.0, atol=2e-5)
    assert np.all(grad <= 0)
    np.testing.assert_allclose(cost, cost_f, rtol=2e-6)
    # one chunk against the oracle (seconds on the CPU)
    b = 17
    off = int(seqlen[:b].sum())
    c_ref, g_ref = oracle.crf_flipflop_loss(scores[:, b:b + 1], seqs[off:off + seqlen[b]],
                                            seqlen[b:b + 1], 1.0, impl='f32')
    np.testing.assert_allclose(cost[b], c_ref[0], rtol=RTOL)
    np.testing.assert_allclose(grad[:, b], g_ref[:, 0], rtol=RTOL, atol=5e-6 / nblk)


def test_host_pointer_dropin_matches_reference_abi(dev, oracle, kat):
    """The (A) layer: reference names + host pointers, called like libctc."""
    import ctypes
    from taiyaki_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    FP, ZP, IP = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_size_t),
                  ctypes.POINTER(ctypes.c_int32))
    nblk, nbatch = 75, 5
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=8)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=9, lengths=[30, 25, 1, 40, 33])
    mv, st = oracle.build_indices(seqs, seqlen, 4)
    sl = seqlen.astype(np.int32)
    score = np.zeros(nbatch, np.float32)
    grad = np.zeros_like(scores)
    lib.crf_flipflop_grad.restype = None
    lib.crf_flipflop_grad.argtypes = [FP, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                      ZP, ZP, IP, FP, FP]
    lib.crf_flipflop_grad(scores.ctypes.data_as(FP), 40, nblk, nbatch, mv.ctypes.data_as(ZP),
                          st.ctypes.data_as(ZP), sl.ctypes.data_as(IP),
                          score.ctypes.data_as(FP), grad.ctypes.data_as(FP))
    impl = 'ref' if oracle.have_ref() else 'f32'
    s_ref, g_ref = oracle.c_crf_flipflop_grad(scores, mv, st, seqlen, impl)
    np.testing.assert_allclose(score, s_ref, rtol=RTOL)
    np.testing.assert_allclose(grad, g_ref, rtol=RTOL, atol=5e-6)
    score2 = np.zeros(nbatch, np.float32)
    lib.crf_flipflop_cost.restype = None
    lib.crf_flipflop_cost.argtypes = [FP, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                      ZP, ZP, IP, FP]
    lib.crf_flipflop_cost(scores.ctypes.data_as(FP), 40, nblk, nbatch, mv.ctypes.data_as(ZP),
                          st.ctypes.data_as(ZP), sl.ctypes.data_as(IP),
                          score2.ctypes.data_as(FP))
    np.testing.assert_allclose(score2, oracle.c_crf_flipflop_cost(scores, mv, st, seqlen, impl),
                               rtol=RTOL)


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'e'])
def test_golden_logz(dev, golden_random, tag):
    from taiyaki_b200 import layers
    g = golden_random
    full = torch.tensor(g[tag + '_scores'], device=dev)
    x = full[:, :, :40]            # strided view for the 45-column cases
    x.requires_grad_(True)
    lz = layers.flipflop_logpartition(x)
    lz.sum().backward()
    np.testing.assert_allclose(lz.detach().cpu().numpy(), g[tag + '_logz'], rtol=2e-6, atol=3e-4)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[tag + '_logz_grad'], rtol=RTOL, atol=3e-6)


@pytest.mark.parametrize('nblk,nbatch', [(1, 2), (7, 3), (100, 5), (800, 64)])
def test_logz_matches_the_reference_gpu_kernels(dev, oracle, nblk, nbatch):
    """csrc/logz.cu against the reference's OWN GPU path on the same device: the CUDA C of its
    CuPy RawKernels (cupy_extensions/flipflop.py:10-296) compiled by oracle/build_cupy_ref.py;
    logZ and its gradient (LogZ.forward / backward, flipflop.py:338-355)."""
    from taiyaki_b200 import layers
    if oracle.libcupy_ref() is None:
        pytest.skip('oracle/_ref/libcupy_ref.so was not built (needs /root/reference at build time)')
    scores = torch.tensor(oracle.synth_scores(nblk, nbatch, 40, seed=nblk), device=dev)
    lz_ref, g_ref = oracle.cupy_ref_logz(scores)
    x = scores.clone().requires_grad_(True)
    lz = layers.flipflop_logpartition(x)
    lz.sum().backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(lz.detach().cpu().numpy(), lz_ref.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g_ref.cpu().numpy(), rtol=RTOL, atol=3e-6)


def test_logz_decodeutil_golden(dev, kat):
    from taiyaki_b200 import layers
    w = torch.tensor(kat['du_weights'][:, None, :], device=dev)
    lz = float(layers.log_partition_flipflop(w))
    assert abs(lz - float(kat['du_logz_flipstart'])) < 3e-5


def test_logz_full_size(dev, oracle):
    from taiyaki_b200 import layers
    nblk, nbatch = 800, 64
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=2)
    x = torch.tensor(scores, device=dev, requires_grad=True)
    lz = layers.flipflop_logpartition(x)
    lz.sum().backward()
    g = x.grad.cpu().numpy()
    np.testing.assert_allclose(g.sum(-1), 1.0, atol=2e-5)
    lz_ref, g_ref = oracle.c_flipflop_logz(scores[:, 5:7], want_grad=True, impl='f64')
    np.testing.assert_allclose(lz.detach().cpu().numpy()[5:7], lz_ref, rtol=2e-6)
    np.testing.assert_allclose(g[:, 5:7], g_ref, rtol=RTOL, atol=3e-6)
    # cost-only path (no gradient requested)
    lz2 = layers.flipflop_logpartition(torch.tensor(scores, device=dev))
    np.testing.assert_allclose(lz2.cpu().numpy(), lz.detach().cpu().numpy(), rtol=1e-6)


def test_training_loss_matches_reference_composition(dev, oracle):
    """loss_b = crf_cost_b + logZ_b / nblk (train_flipflop.py:166-182) and its
    gradient, against the oracle."""
    from taiyaki_b200 import ctc, layers
    nblk, nbatch = 100, 6
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=21)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, seed=22)
    x = torch.tensor(scores, device=dev, requires_grad=True)
    lossvec = ctc.crf_flipflop_loss(x, torch.tensor(seqs), torch.tensor(seqlen), 1.0)
    lossvec = lossvec + layers.flipflop_logpartition(x) / nblk
    lossvec.mean().backward()
    c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='f32')
    lz_ref, gz_ref = oracle.c_flipflop_logz(scores, want_grad=True, impl='f32')
    np.testing.assert_allclose(lossvec.detach().cpu().numpy(), c_ref + lz_ref / nblk, rtol=RTOL,
                               atol=1e-5)
    # a difference of two posteriors: each term within 1e-4 of its own magnitude
    err = np.abs(x.grad.cpu().numpy() * nbatch - (g_ref + gz_ref / nblk))
    assert np.all(err <= RTOL * (np.abs(g_ref) + np.abs(gz_ref) / nblk) + 1e-5 / nblk)


@pytest.mark.parametrize('ntrans', [40, 45])
def test_fused_train_loss_equals_separate_operators(dev, oracle, ntrans):
    """ctc.flipflop_train_loss == crf/cat-mod loss + logZ/nblk (values and gradient),
    and both match the oracle composition."""
    from taiyaki_b200 import ctc, layers
    nblk, nbatch = 150, 7
    scores = oracle.synth_scores(nblk, nbatch, ntrans, seed=31)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, nbatch, seed=32)
    kw = {}
    if ntrans == 45:
        mod_cats = np.concatenate([(r == 1).astype(np.int64) * (np.arange(len(r)) % 2)
                                   for r in raw])
        off = np.array([0, 1, 3, 4, 5], dtype=np.int32)
        w = np.array([1.0, 1.0, 0.6, 1.0, 1.0], dtype=np.float32)
        kw = dict(mod_cats=torch.tensor(mod_cats), can_mods_offsets=off, mod_cat_weights=w)
    st, sl = torch.tensor(seqs), torch.tensor(seqlen)
    x1 = torch.tensor(scores, device=dev, requires_grad=True)
    l1 = ctc.flipflop_train_loss(x1, st, sl, 1.5, **kw)
    l1.mean().backward()
    x2 = torch.tensor(scores, device=dev, requires_grad=True)
    if ntrans == 45:
        l2 = ctc.cat_mod_flipflop_loss(x2, st, sl, kw['mod_cats'], off, w, 1.5)
    else:
        l2 = ctc.crf_flipflop_loss(x2, st, sl, 1.5)
    l2 = l2 + layers.flipflop_logpartition(x2[:, :, :40]) / nblk
    l2.mean().backward()
    np.testing.assert_allclose(l1.detach().cpu().numpy(), l2.detach().cpu().numpy(), rtol=1e-6,
                               atol=1e-6)
    np.testing.assert_allclose(x1.grad.cpu().numpy(), x2.grad.cpu().numpy(), rtol=1e-5, atol=1e-9)
    if ntrans == 45:
        c_ref, g_ref = oracle.cat_mod_flipflop_loss(scores, seqs, seqlen, mod_cats, off, w, 1.5,
                                                    impl='f32')
    else:
        c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.5, impl='f32')
    lz, gz = oracle.c_flipflop_logz(np.ascontiguousarray(scores[:, :, :40]), impl='f32')
    np.testing.assert_allclose(l1.detach().cpu().numpy(), c_ref + lz / nblk, rtol=RTOL, atol=1e-5)
    want = g_ref.copy()
    want[:, :, :40] += gz / nblk
    tol = RTOL * np.abs(g_ref)
    tol[:, :, :40] += RTOL * np.abs(gz) / nblk
    err = np.abs(x1.grad.cpu().numpy() * nbatch - want)
    assert np.all(err <= tol + 1e-5 / nblk), (err - tol).max() * nblk


def test_bad_label_raises_reference_assertion_lazily(dev, oracle):
    """ctc.pyx:133-134 asserts label indices on the host; here the kernels clamp and a
    device flag is raised when somebody collects it (no synchronisation per call)."""
    from taiyaki_b200 import ctc
    ctc.check_pending()
    scores = torch.tensor(oracle.synth_scores(30, 2, 40, seed=1), device=dev)
    seqs = torch.tensor([0, 5, 1, 9, 2, 6])       # 9 is not a flip-flop label for 4 bases
    cost = ctc.crf_flipflop_loss(scores, seqs, torch.tensor([3, 3]), 1.0)
    assert torch.isfinite(cost).all()             # clamped, nothing read out of bounds
    with pytest.raises(AssertionError, match='out of range'):
        ctc.check_pending()
    ctc.crf_flipflop_loss(scores, torch.tensor([0, 5, 1, 7, 2, 6]), torch.tensor([3, 3]), 1.0)
    ctc.check_pending()                           # valid labels: no flag
