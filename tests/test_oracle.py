"""Pin the CPU oracle (oracle/) to the reference: its embedded known-answer
tables, golden vectors generated from the reference (tests/golden/
make_golden.py) and, when present, the reference's own C (oracle/_ref)."""
import numpy as np
import pytest

CASES40 = ['a', 'b', 'e']
CASES45 = ['c', 'd']


def test_kat_crf_twostate(oracle, kat):
    # c_crf_flipflop.c:520-695 -> "Forwards scores: -2.378088 -2.378088"
    for impl in ('f32', 'f64'):
        fb = oracle.c_scores_fb(kat['crf_logprob'], kat['crf_move'], kat['crf_stay'],
                                kat['crf_seqlen'], impl=impl)
        np.testing.assert_allclose(fb, -2.378088, atol=2e-6)
        sc, gr = oracle.c_crf_flipflop_grad(kat['crf_logprob'], kat['crf_move'],
                                            kat['crf_stay'], kat['crf_seqlen'], impl)
        np.testing.assert_allclose(sc, kat['crf_score'], atol=2e-6)
        np.testing.assert_allclose(gr, kat['crf_grad'], atol=2e-6)
        np.testing.assert_allclose(gr[:, 0], gr[:, 1], atol=1e-7)
        np.testing.assert_allclose(gr.sum(-1), 1.0, atol=1e-5)


def test_kat_cat_mod(oracle, kat):
    # c_cat_mod_flipflop.c:586-870 -> "-52.354622 -195.435257"
    np.testing.assert_allclose(kat['cm_score'], [-52.354622, -195.435257], rtol=1e-6)
    for impl in ('f32', 'f64'):
        fb = oracle.c_scores_fb(kat['cm_logprob'], kat['cm_move'], kat['cm_stay'],
                                kat['cm_seqlen'], kat['cm_modmove'], kat['cm_modfact'],
                                impl=impl)
        np.testing.assert_allclose(fb[:, 0], [-52.354622, -195.435257], rtol=1e-6)
        np.testing.assert_allclose(fb[:, 1], [-52.354622, -195.435257], rtol=1e-6)
        sc, gr = oracle.c_cat_mod_flipflop_grad(
            kat['cm_logprob'], kat['cm_move'], kat['cm_stay'], kat['cm_modmove'],
            kat['cm_modfact'], kat['cm_seqlen'], impl)
        np.testing.assert_allclose(sc, kat['cm_score'], rtol=1e-6)
        np.testing.assert_allclose(gr, kat['cm_grad'], rtol=1e-4, atol=1e-5)


def test_unit_ctc_loss_path_probabilities(oracle, kat):
    # test/unit/test_ctc_loss.py:84-103
    scores = kat['unit_scores']
    lz = oracle.flipflop_logpartition(scores)
    assert abs(float(lz[0])) < 1e-6
    assert abs(float(kat['unit_logpart'])) < 1e-6
    for seq, prob in zip(kat['unit_seqs'], kat['unit_probs']):
        for impl in ('f32', 'ref') if oracle.have_ref() else ('f32',):
            cost = oracle.crf_flipflop_loss(scores, seq, [3], 1.0, want_grad=False, impl=impl)
            assert abs(float(np.exp(-cost[0] * 4)) - prob) < 1e-7


def test_unit_ctc_loss_gradient_check(oracle, kat):
    # test/unit/test_ctc_loss.py:105-135
    scores = kat['unit_scores']
    rng = np.random.RandomState(0)
    for seq in kat['unit_seqs'][:2]:
        cost, grad = oracle.crf_flipflop_loss(scores, seq, [3], 1.0, impl='f32')
        dx = (rng.standard_normal(scores.shape) * 1e-3).astype(np.float32)
        cost2 = oracle.crf_flipflop_loss(scores + dx, seq, [3], 1.0, want_grad=False, impl='f32')
        est = float((dx * grad).sum())
        assert abs((cost2[0] - cost[0]) / cost[0] - est / cost[0]) < 1e-5


def test_decodeutil_logz_golden(oracle, kat):
    # test/unit/test_decodeutil.py:16-31: flip-start partition function
    w = kat['du_weights'][:, None, :]
    for impl in ('f32', 'f64'):
        lz = oracle.c_flipflop_logz(w, want_grad=False, impl=impl)
        assert abs(float(lz[0]) - float(kat['du_logz_flipstart'])) < 2e-5
    lz2 = oracle.c_flipflop_logz(w, want_grad=False, flop_init=-1e30)
    assert abs(float(lz2[0]) - float(kat['du_logz_flipstart'])) < 2e-5


def test_flipflop_code_golden(oracle, kat):
    np.testing.assert_array_equal(oracle.flipflop_code(kat['code_in0']), kat['code_out0'])
    np.testing.assert_array_equal(kat['code_out0'], [1, 3, 2, 3, 7, 3, 7, 1, 5])
    c = oracle.flipflop_code(kat['code_in1'])
    np.testing.assert_array_equal(c, kat['code_out1'])
    np.testing.assert_array_equal(oracle.move_indices(c), kat['code_move1'])
    np.testing.assert_array_equal(oracle.stay_indices(c), kat['code_stay1'])


@pytest.mark.parametrize('tag', CASES40)
def test_random_crf_vs_reference_golden(oracle, golden_random, tag):
    g = golden_random
    scores, seqs, seqlen = g[tag + '_scores'], g[tag + '_seqs'], g[tag + '_seqlen']
    sharp = float(g[tag + '_sharp'])
    mv, st = oracle.build_indices(seqs, seqlen, 4)
    np.testing.assert_array_equal(mv, g[tag + '_move'])
    np.testing.assert_array_equal(st, g[tag + '_stay'])
    lp = np.float32(sharp) * scores
    for impl in ('f32', 'f64'):
        sc, gr = oracle.c_crf_flipflop_grad(lp, mv, st, seqlen, impl)
        np.testing.assert_allclose(sc, g[tag + '_score'], rtol=2e-5, atol=1e-4)
        np.testing.assert_allclose(gr, g[tag + '_grad'], rtol=1e-4, atol=2e-6)
        sc2 = oracle.c_crf_flipflop_cost(lp, mv, st, seqlen, impl)
        np.testing.assert_allclose(sc2, g[tag + '_score_costonly'], rtol=2e-5, atol=1e-4)
    nblk = scores.shape[0]
    if tag + '_torch_flipfloploss' in g.files:
        # taiyaki/loss.py FlipFlopLoss forward == -score/nblk/sharp
        cost = oracle.crf_flipflop_loss(scores, seqs, seqlen, sharp, want_grad=False, impl='f32')
        np.testing.assert_allclose(cost, g[tag + '_torch_flipfloploss'].ravel(), rtol=2e-5)
    # invariants: rows of G sum to one for non-empty chunks, zero for empty
    _, gr = oracle.c_crf_flipflop_grad(lp, mv, st, seqlen, 'f32')
    rows = gr.sum(-1)
    for b, L in enumerate(seqlen):
        np.testing.assert_allclose(rows[:, b], 1.0 if L > 0 else 0.0, atol=2e-5)
    assert nblk == rows.shape[0]


@pytest.mark.parametrize('tag', CASES45)
def test_random_cat_mod_vs_reference_golden(oracle, golden_random, tag):
    g = golden_random
    scores, seqs, seqlen = g[tag + '_scores'], g[tag + '_seqs'], g[tag + '_seqlen']
    sharp = float(g[tag + '_sharp'])
    mm, mf = oracle.build_mod_indices(seqs, seqlen, g[tag + '_mod_cats'],
                                      g[tag + '_can_mods_offsets'],
                                      g[tag + '_mod_cat_weights'], 4)
    np.testing.assert_array_equal(mm, g[tag + '_modmove'])
    np.testing.assert_array_equal(mf, g[tag + '_modfact'])
    ts = np.ones(45, dtype=np.float32)
    ts[:40] = sharp
    lp = np.ascontiguousarray(scores * ts)
    for impl in ('f32', 'f64'):
        sc, gr = oracle.c_cat_mod_flipflop_grad(lp, g[tag + '_move'], g[tag + '_stay'],
                                                mm, mf, seqlen, impl)
        np.testing.assert_allclose(sc, g[tag + '_score'], rtol=2e-5, atol=1e-4)
        np.testing.assert_allclose(gr, g[tag + '_grad'], rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize('tag', CASES40 + CASES45)
def test_random_logz_vs_reference_torchscript(oracle, golden_random, tag):
    g = golden_random
    w = np.ascontiguousarray(g[tag + '_scores'][:, :, :40])
    for impl in ('f32', 'f64'):
        lz, gr = oracle.c_flipflop_logz(w, want_grad=True, impl=impl)
        np.testing.assert_allclose(lz, g[tag + '_logz'], rtol=2e-6, atol=2e-4)
        np.testing.assert_allclose(gr, g[tag + '_logz_grad'], rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(gr.sum(-1), 1.0, atol=2e-5)


def test_restatement_vs_reference_c_live(oracle):
    """When the reference's C is built (oracle/_ref), compare live on a seeded
    ragged batch, including sharpening identity loss(x,2) == loss(2x,1)/2."""
    if not oracle.have_ref():
        pytest.skip('oracle/_ref not built')
    nblk, nbatch = 96, 6
    scores = oracle.synth_scores(nblk, nbatch, 40, seed=3)
    seqs, seqlen, _ = oracle.synth_seqs(nblk, nbatch, stride=5, seed=4,
                                        lengths=[40, 1, 52, 33, 0, 47])
    c_ref, g_ref = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='ref')
    c_f32, g_f32 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='f32')
    c_f64, g_f64 = oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, impl='f64')
    np.testing.assert_allclose(c_f32, c_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g_f32, g_ref, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(c_f64, c_ref, rtol=1e-5, atol=1e-6)
    a = oracle.crf_flipflop_loss(scores, seqs, seqlen, 2.0, want_grad=False, impl='ref')
    b = oracle.crf_flipflop_loss(2 * scores, seqs, seqlen, 1.0, want_grad=False, impl='ref')
    np.testing.assert_allclose(a, b / 2, rtol=1e-6)
    assert np.all(g_ref <= 0)


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_decode_golden_vs_restatement(oracle, tag):
    """tests/golden/decode.npz (the reference's PyTorch Viterbi and make_trans) against the
    numpy Viterbi restatement and the C partition-function gradient."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'decode.npz'))
    scores = g[tag + '_scores']
    fwd, tb, path = oracle.flipflop_viterbi(scores)
    np.testing.assert_array_equal(path, g[tag + '_path'])
    np.testing.assert_array_equal(tb, g[tag + '_tb'])
    np.testing.assert_allclose(fwd, g[tag + '_fwd'], rtol=1e-6, atol=1e-5)
    _, trans = oracle.c_flipflop_logz(np.ascontiguousarray(scores), want_grad=True, impl='f64')
    np.testing.assert_allclose(trans, g[tag + '_trans'], rtol=2e-4, atol=2e-6)
