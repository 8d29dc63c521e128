"""Basecalling driver (taiyaki_b200/basecall_helpers.py, qscores.py, basecall.py;
SURVEY 8(f) row 3) against tests/golden/basecall.npz, which make_golden.py writes from
the reference's own basecall_helpers / qscores / decode / path_to_str.

CPU tests: chunking, stitching, error probabilities and quality strings are bit-exact
(index arithmetic and a 40 x 4 product of the same floats).  GPU tests: the flow after
the network (posterior weights -> Viterbi -> stitch -> bases + qualities) on the golden
scores, and the whole driver on a random-weight network for internal consistency."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'basecall.npz')
CASES = 'abcdefg'


@pytest.fixture(scope='module')
def g():
    return np.load(GOLDEN)


@pytest.mark.parametrize('tag', CASES)
@pytest.mark.parametrize('as_tensor', [False, True])
def test_chunk_read_golden(g, tag, as_tensor):
    from taiyaki_b200 import basecall_helpers
    nsample, chunk_size, overlap, stride = g[tag + '_cfg']
    signal = g[tag + '_signal']
    if as_tensor:
        signal = torch.tensor(signal)
    chunks, cs, ce = basecall_helpers.chunk_read(signal, int(chunk_size), int(overlap))
    if as_tensor:
        chunks = chunks.numpy()
    assert chunks.dtype == np.float32
    np.testing.assert_array_equal(chunks, g[tag + '_chunks'])
    np.testing.assert_array_equal(cs, g[tag + '_starts'])
    np.testing.assert_array_equal(ce, g[tag + '_ends'])


@pytest.mark.parametrize('tag', CASES)
def test_stitch_chunks_golden(g, tag):
    from taiyaki_b200 import basecall_helpers
    stride = int(g[tag + '_cfg'][3])
    cs, ce = g[tag + '_starts'], g[tag + '_ends']
    out = torch.tensor(g[tag + '_out'])
    path = torch.tensor(g[tag + '_path'].astype(np.int64))
    np.testing.assert_array_equal(
        basecall_helpers.stitch_chunks(out, cs, ce, stride).numpy(), g[tag + '_stitched'])
    np.testing.assert_array_equal(
        basecall_helpers.stitch_chunks(path, cs, ce, stride).numpy(), g[tag + '_stitched_path'])
    np.testing.assert_array_equal(
        basecall_helpers.stitch_chunks(path, cs, ce, stride, path_stitching=True).numpy(),
        g[tag + '_stitched_path_ps'])


def test_stitched_length_covers_read(g):
    """Size-independent property: the stitched output of a long read has one block per
    `stride` samples (up to the rounding of the halved overlaps)."""
    from taiyaki_b200 import basecall_helpers
    for nsample, chunk_size, overlap, stride in ((1234567, 5000, 500, 5), (400001, 2000, 200, 2)):
        cs, ce = basecall_helpers.chunk_bounds(nsample, chunk_size, overlap)
        assert cs[0] == 0 and ce[-1] == nsample and np.all(ce - cs == chunk_size)
        start, end = basecall_helpers.stitch_ranges(cs, ce, stride)
        kept = int((end - start).sum())
        assert abs(kept - nsample // stride) <= len(cs)
        # ranges of neighbouring chunks meet: same absolute block, no gap, no overlap
        absolute_end = cs[:-1] // stride + end[:-1]
        absolute_start = cs[1:] // stride + start[1:]
        assert np.all(np.abs(absolute_end - absolute_start) <= 1)


def test_errprobs_and_qstring_golden(g):
    from taiyaki_b200 import qscores
    trans = torch.tensor(g['q_trans'])
    paths = torch.tensor(g['q_paths'].astype(np.int64))
    ep = qscores.errprobs_from_trans(trans, paths)
    np.testing.assert_allclose(ep.numpy(), g['q_errprobs'], rtol=0, atol=2e-7)
    assert bool((ep[0] == -1.0).all())
    assert qscores.qchar_from_errprob(np.array([0.5, 0.1, 0.011, 1e-4, 0.9999]), 1.0, 0.0) == str(g['q_qchars'])
    assert qscores.qchar_from_errprob(np.array([0.5, 0.1, 0.011, 1e-4]), 0.9, 1.5) == str(g['q_qchars_cal'])
    for b in range(4):      # the constant matrix marks exactly transitions_into_base
        m = qscores._into_base_matrix(4, 'cpu')[:, b]
        assert sorted(torch.nonzero(m).flatten().tolist()) == sorted(
            qscores.transitions_into_base(b, 4, 'cpu').tolist())


@pytest.mark.parametrize('tag', ['post', 'temp'])
def test_flow_from_golden_paths_cpu(g, tag):
    """Host half of the flow: the reference's per-chunk paths and decoded weights ->
    stitched path, bases, error probabilities, quality string."""
    from taiyaki_b200 import basecall_helpers, qscores
    from taiyaki_b200.flipflopfings import path_to_str
    stride = int(g['flow_cfg'][3])
    cs, ce = g['flow_starts'], g['flow_ends']
    paths = torch.tensor(g['flow_%s_chunk_paths' % tag].astype(np.int64))
    best = basecall_helpers.stitch_chunks(paths, cs, ce, stride).numpy()
    np.testing.assert_array_equal(best, g['flow_%s_path' % tag])
    assert path_to_str(best, alphabet='ACGT', include_first_source=False) == str(g['flow_%s_basecall' % tag])
    if tag == 'post':
        trans = torch.tensor(g['flow_post_trans'])
        ep = basecall_helpers.stitch_chunks(qscores.errprobs_from_trans(trans, paths), cs, ce, stride)
        np.testing.assert_allclose(ep.numpy(), g['flow_post_errprobs'], rtol=0, atol=2e-6)
        q = qscores.path_errprobs_to_qstring(ep, best, 1.0, 0.0)
        ref_q = str(g['flow_post_qstring'])
        assert len(q) == len(ref_q)
        # a quality character can differ by one where the score sits on a rounding edge
        assert sum(a != b for a, b in zip(q, ref_q)) <= 2
        assert max(abs(ord(a) - ord(b)) for a, b in zip(q, ref_q)) <= 1


def test_basecall_cli_host_pieces(tmp_path):
    """bin/basecall.py: the reference's flags and defaults (bin/basecall.py:23-72), signal
    iteration from a folder of .npy files or one .npz, --limit and a strand list."""
    import importlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bin'))
    cli = importlib.import_module('basecall')
    a = cli.get_parser().parse_args(['reads', 'model.checkpoint'])
    assert (a.chunk_size, a.overlap, a.max_concurrent_chunks, a.posterior, a.fastq, a.temperature,
            a.qscore_scale, a.qscore_offset, a.reverse, a.alphabet) == (
                1000, 100, 128, True, False, 1.0, 1.0, 0.0, False, 'ACGT')
    a = cli.get_parser().parse_args(['--fastq', '--no-posterior', '--chunk_size', '500', 'r', 'm'])
    assert a.fastq is True and a.posterior is False and a.chunk_size == 500
    folder = tmp_path / 'sig'
    folder.mkdir()
    for i in range(3):
        np.save(str(folder / ('r%d.npy' % i)), np.arange(10 + i, dtype='f4'))
    (folder / 'notes.txt').write_text('x')
    got = list(cli.iterate_signals(str(folder)))
    assert [g[0] for g in got] == ['r0', 'r1', 'r2'] and len(got[2][1]) == 12
    assert [g[0] for g in cli.iterate_signals(str(folder), limit=2)] == ['r0', 'r1']
    strands = tmp_path / 's.tsv'
    strands.write_text('filename\tread_id\na\tr2\n')
    assert [g[0] for g in cli.iterate_signals(str(folder), strand_list=str(strands))] == ['r2']
    np.savez(str(tmp_path / 'all.npz'), x=np.zeros(4), y=np.ones(5))
    assert [(k, len(v)) for k, v in cli.iterate_signals(str(tmp_path / 'all.npz'))] == [('x', 4), ('y', 5)]


def _normalisation_cases():
    rng = np.random.RandomState(0)
    for n in (62001, 62000, 5, 2, 1001):
        x = 90 + 12 * rng.standard_normal(n)
        x[::7] = np.round(x[::7])               # ties around the median
        yield x


@pytest.mark.parametrize('device', ['cpu', pytest.param('cuda:0', marks=pytest.mark.gpu)])
def test_normalisation_on_device_is_numpy_s(device):
    """Median / MAD normalisation by sorting on the device against the numpy code of
    bin/basecall.py:74-87 on float64 signals: bit-identical, odd and even lengths, reversed,
    and with per-read shift / scale."""
    from taiyaki_b200 import basecall
    for x in _normalisation_cases():
        for rev in (False, True):
            want = basecall.normalise_signal(x, rev, None)
            got = basecall.normalise_signal_device(x, device, rev, None)
            assert got.dtype == torch.float32 and want.dtype == np.float32
            np.testing.assert_array_equal(got.cpu().numpy(), want)
        p = {'shift': 83.7, 'scale': 15.28}
        np.testing.assert_array_equal(
            basecall.normalise_signal_device(x, device, True, p).cpu().numpy(),
            basecall.normalise_signal(x, True, p))


# ------------------------------------------------------------------ GPU

@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


def _close(a, b, ratio):
    """Equal, or near-equal strings: the projection GEMMs may pick another algorithm for
    another batch size, and a last-bit difference can move a bf16 rounding of the
    recurrent state, so calls of the same read in different batches may differ at
    isolated positions."""
    import difflib
    if a == b:
        return True
    return difflib.SequenceMatcher(None, a, b, autojunk=False).ratio() >= ratio


def _path_score(trans, path):
    """Score of a flip-flop path under [T, S] transition weights (fp64)."""
    frm, to = path[:-1], path[1:]
    idx = np.where(to < 4, to * 8 + frm, 32 + frm)
    return float(trans[np.arange(len(idx)), idx].astype(np.float64).sum())


@pytest.mark.gpu
@pytest.mark.parametrize('tag,posterior,temperature', [('raw', False, 1.0), ('post', True, 1.0),
                                                      ('temp', True, 0.5)])
def test_decode_flow_golden(g, dev, tag, posterior, temperature):
    """Scores of one read's chunks -> bases: the device flow against the reference's.
    Without --posterior the Viterbi path is bit-exact.  With it the weights pass through
    the partition-function kernels (rtol 1e-4) and a logarithm before Viterbi, so a near
    tie may resolve differently: each chunk's path must score within 0.05 of the
    reference's best under the reference's weights (0.05 = 2 x 200 blocks x 1e-4), and the basecall must match the
    reference's except for isolated positions."""
    from taiyaki_b200 import basecall, basecall_helpers
    from taiyaki_b200.flipflopfings import path_to_str
    stride = int(g['flow_cfg'][3])
    cs, ce = g['flow_starts'], g['flow_ends']
    scores = torch.tensor(g['flow_scores'], device=dev)
    trans, paths = basecall.decode_chunks(scores, posterior, temperature)
    torch.cuda.synchronize()
    ref_paths = g['flow_%s_chunk_paths' % tag].astype(np.int64)
    got_paths = paths.cpu().numpy()
    if not posterior:
        np.testing.assert_array_equal(got_paths, ref_paths)
    else:
        if tag == 'post':
            # compared as weights (the bar of test_make_trans_golden): the logarithm of a
            # weight near the 1e-8 floor magnifies an absolute error that no path uses
            np.testing.assert_allclose(np.exp(trans.cpu().numpy()), np.exp(g['flow_post_trans']),
                                       rtol=1e-4, atol=1e-6)
            w = g['flow_post_trans']
            for c in range(ref_paths.shape[1]):
                assert _path_score(w[:, c], got_paths[:, c]) >= _path_score(w[:, c], ref_paths[:, c]) - 0.05
        assert (got_paths != ref_paths).mean() < 0.03
    call, qstring = basecall._finish_read(trans, paths, cs, ce, stride, 'ACGT', posterior, 1.0, 0.0)
    ref_call = str(g['flow_%s_basecall' % tag])
    if np.array_equal(got_paths, ref_paths):
        assert call == ref_call
        if posterior:
            ref_q = str(g['flow_%s_qstring' % tag])
            assert len(qstring) == len(ref_q) == len(call)
            assert max(abs(ord(a) - ord(b)) for a, b in zip(qstring, ref_q)) <= 1
    else:
        assert abs(len(call) - len(ref_call)) <= 0.01 * len(ref_call) + 2
    best = basecall_helpers.stitch_chunks(paths, cs, ce, stride).cpu().numpy()
    assert path_to_str(best, alphabet='ACGT', include_first_source=False) == call


@pytest.mark.gpu
def test_driver_end_to_end(dev, tmp_path):
    """process_signal / process_signals / run_model / bin/basecall.py on a random-weight
    mLstm_flipflop: pooled batches give the same calls as reads called alone, the
    megalodon hook returns the stitched network output, the entry point writes fastq."""
    import sys
    from taiyaki_b200 import basecall, basecall_helpers, helpers
    from taiyaki_b200.alphabet import AlphabetInfo
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    torch.manual_seed(3)
    ai = AlphabetInfo('ACGT', 'ACGT')
    model = helpers.load_model(os.path.join(root, 'models', 'mLstm_flipflop.py'),
                               model_metadata={'reverse': False, 'standardize': True}, stride=5,
                               winlen=19, insize=1, size=64, alphabet_info=ai).to(dev)
    stride = helpers.guess_model_stride(model)
    assert stride == 5
    rng = np.random.RandomState(5)
    signals = [('r%d' % i, (90 + 12 * rng.standard_normal(n)).astype('f4'))
               for i, n in enumerate((7000, 1999, 2000, 12345, 600))]
    signals.append(('missing', None))
    chunk, ovl = 400 * stride, 40 * stride
    alone = [(rid, *basecall.process_signal(s, model, chunk, ovl, None, 40, stride, 'ACGT', 4,
                                            fastq=True)) for rid, s in signals]
    pooled = basecall.process_signals(signals, model, chunk, ovl, {}, 40, stride, 'ACGT', 4, fastq=True)
    assert [(r[0], r[3]) for r in pooled] == [(r[0], r[3]) for r in alone]
    for a, b in zip(pooled[:-1], alone[:-1]):
        assert _close(a[1], b[1], 0.97) and len(a[2]) == len(a[1]) and len(b[2]) == len(b[1])
    for rid, call, q, nsample in pooled[:-1]:
        assert nsample > 0 and set(call) <= set('ACGT') and len(q) == len(call)
    assert pooled[-1] == ('missing', None, None, 0)
    # per-read scaling in place of med/MAD normalisation
    from taiyaki_b200.maths import med_mad
    med, mad = med_mad(signals[0][1])
    scaled = basecall.process_signal(signals[0][1], model, chunk, ovl, {'shift': med, 'scale': mad},
                                     40, stride, 'ACGT', 4, fastq=True)
    assert _close(scaled[0], alone[0][1], 0.97)
    with pytest.raises(NotImplementedError):
        basecall.process_signal(signals[0][1], model, chunk, ovl, None, 40, stride, 'ACGT', 4, beam=(5, False))
    # megalodon hook: [blocks, 40] for the whole read, equal to chunk-by-chunk stitching
    normed = basecall.med_mad_norm(signals[3][1])
    out = basecall_helpers.run_model(normed, model, 400, 40, max_concur_chunks=3)
    assert abs(out.shape[0] - len(normed) // stride) <= 2 and out.shape[1] == 40 and np.isfinite(out).all()
    out_all = basecall_helpers.run_model(normed, model, 400, 40, return_numpy=False)
    assert out_all.device.type == 'cuda'
    assert out_all.shape == out.shape and float(np.abs(out_all.cpu().numpy() - out).mean()) < 1e-2
    # entry point
    sys.path.insert(0, os.path.join(root, 'bin'))
    import importlib
    bc = importlib.import_module('basecall')
    folder = tmp_path / 'reads'
    folder.mkdir()
    for rid, s in signals[:-1]:
        np.save(str(folder / (rid + '.npy')), s)
    ckpt, _ = helpers.save_model(model, str(tmp_path))
    outfile = tmp_path / 'calls.fq'
    bc.main(['--chunk_size', '400', '--overlap', '40', '--max_concurrent_chunks', '4', '--fastq',
             '--output', str(outfile), '--reads_per_batch', '3', str(folder), ckpt])
    lines = outfile.read_text().splitlines()
    assert len(lines) == 4 * sum(1 for r in pooled[:-1] if len(r[1]) > 0)   # empty calls are not written
    called = {lines[i][1:]: (lines[i + 1], lines[i + 3]) for i in range(0, len(lines), 4)}
    for rid, call, q, _ in pooled[:-1]:
        if rid in called:
            assert _close(called[rid][0], call, 0.97) and len(called[rid][1]) == len(called[rid][0])
