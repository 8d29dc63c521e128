"""Remapping DP (csrc/remap.cu, taiyaki_b200/flipflop_remap.py) against
tests/golden/remap.npz -- alignments of the reference's taiyaki/flipflop_remap.py
(make_golden.py remap) including the two tables of its own unit test.  The DP is adds and
maxima of fp32 scores promoted to fp64: scores and paths are compared bit-exact."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'remap.npz')
CASES = 'abcdefghijk'


@pytest.fixture(scope='module')
def g():
    return np.load(GOLDEN)


def test_oracle_matches_reference(g):
    from oracle import oracle
    for tag in CASES:
        bases = np.array(['ACGT'.find(b) for b in str(g[tag + '_seq'])])
        step, stay = oracle.remap_indices(bases, 4)
        score, path = oracle.map_to_crf_viterbi(g[tag + '_scores'], step, stay, float(g[tag + '_localpen']))
        assert score == float(g[tag + '_score'])
        np.testing.assert_array_equal(path, g[tag + '_path'])
    for tag in ('kat1', 'kat2'):       # test/unit/test_flipflop_remap.py:8-90
        bases = np.array(['AB'.find(b) for b in str(g[tag + '_seq'])])
        step, stay = oracle.remap_indices(bases, 2)
        score, path = oracle.map_to_crf_viterbi(g[tag + '_scores'], step, stay, -0.5)
        assert score == float(g[tag + '_score'])
        np.testing.assert_array_equal(path, g[tag + '_path'])
    assert float(g['kat1_score']) == 6.0 and float(g['kat2_score']) == 3.5


def test_remap_indices_mirror(g):
    from oracle import oracle
    from taiyaki_b200 import flipflop_remap
    for tag in CASES:
        seq = str(g[tag + '_seq'])
        step, stay = flipflop_remap.remap_indices(seq)
        ostep, ostay = oracle.remap_indices(np.array(['ACGT'.find(b) for b in seq]), 4)
        np.testing.assert_array_equal(step, ostep)
        np.testing.assert_array_equal(stay, ostay)
    step, stay = flipflop_remap.remap_indices('AABA', 'AB')       # the unit test's low-level indices
    assert list(step) == [8, 6, 1] and list(stay) == [0, 10, 5, 0]


def test_path_to_ref_to_signal_matches_reference(g):
    """Block path -> Ref_to_signal (from_remapping_path / get_reftosignal) against the
    reference's get_reftosignal on the golden paths: global, clipped at either end."""
    from taiyaki_b200.signal_mapping import SignalMapping
    from taiyaki_b200.mapped_signal_files import check_read
    nclipped = 0
    for tag in CASES:
        stride, signalstart, nd = (int(x) for x in g[tag + '_reftosig_cfg'])
        seq = str(g[tag + '_seq'])
        ref = SignalMapping.get_integer_reference(seq, 'ACGT')
        sm = SignalMapping.from_remapping_path(g[tag + '_path'], ref, stride,
                                               np.zeros(nd, dtype=np.int16), signalstart, read_id=tag)
        assert sm.Ref_to_signal.dtype == np.int32
        np.testing.assert_array_equal(sm.Ref_to_signal, g[tag + '_reftosig'])
        assert check_read(sm) == 'pass'
        nclipped += int((g[tag + '_path'] == -1).any())
        d = sm.get_read_dictionary()
        assert sorted(d) == sorted(['Dacs', 'Ref_to_signal', 'Reference', 'read_id', 'shift_frompA',
                                    'scale_frompA', 'range', 'offset', 'digitisation'])
    assert nclipped >= 3
    # a fully clipped read maps nothing (signal_mapping.py:241-243)
    np.testing.assert_array_equal(SignalMapping.get_reftosignal(np.full(10, -1), 3, 10), [-1] * 4)


def test_launch_groups():
    from taiyaki_b200.flipflop_remap import launch_groups
    assert launch_groups([3, 3, 3], cap=10) == [[0, 1, 2]]
    assert launch_groups([6, 6, 3, 20, 1, 1], cap=10) == [[0], [1, 2], [3], [4, 5]]
    assert launch_groups([], cap=10) == []
    assert launch_groups([8000 * 3500] * 148) == [list(range(148))]       # the benchmark batch: one launch


def _write_remap_inputs(tmp_path, with_mods=True):
    """Three raw reads (.npz), their per-read parameters, references and a strand list."""
    rng = np.random.RandomState(9)
    folder = tmp_path / 'raw'
    folder.mkdir()
    tsv = ['UUID\ttrim_start\ttrim_end\tshift\tscale']
    fasta, reads = [], {}
    for i, (n, L) in enumerate(((6000, 500), (3011, 200), (4500, 60))):
        rid = 'read%d' % i
        letters = 'ACGTZ' if with_mods else 'ACGT'
        ref = ''.join(letters[b] for b in rng.randint(0, len(letters), size=L))
        dacs = rng.randint(300, 700, size=n).astype(np.int16)
        np.savez(str(folder / (rid + '.npz')), dacs=dacs, offset=10.0, range=1400.0, digitisation=8192.0)
        tsv.append('%s\t%d\t20\t85.0\t14.0' % (rid, 50 * i))
        fasta += ['>%s some description' % rid, ref[:70], ref[70:]]
        reads[rid] = (dacs, ref)
    np.savez(str(folder / 'orphan.npz'), dacs=reads['read0'][0], offset=0.0, range=1.0, digitisation=1.0)
    (tmp_path / 'params.tsv').write_text('\n'.join(tsv + ['broken\tline']) + '\n')
    (tmp_path / 'refs.fa').write_text('\n'.join(fasta) + '\n')
    return folder, tmp_path / 'params.tsv', tmp_path / 'refs.fa', reads


def _load_cli(name):
    import importlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bin'))
    return importlib.import_module(name)


def test_prepare_mapped_reads_host_pieces(tmp_path):
    """Inputs of bin/prepare_mapped_reads.py: per-read parameter table, fasta references,
    raw read iteration, the alphabet built from --mod (sorted like the reference's
    do_reorder=True) and the trimming rule of taiyaki/signal.py:77-95."""
    from taiyaki_b200 import prepare_mapping_funcs as pmf
    cli = _load_cli('prepare_mapped_reads')
    folder, tsv, fa, reads = _write_remap_inputs(tmp_path)
    params = pmf.get_per_read_params_dict_from_tsv(str(tsv))
    assert sorted(params) == ['read0', 'read1', 'read2']
    assert params['read2'] == {'trim_start': 100, 'trim_end': 20, 'shift': 85.0, 'scale': 14.0}
    refs = pmf.fasta_file_to_dict(str(fa), alphabet='ACGTZ')
    assert {k: v for k, v in refs.items()} == {k: v[1] for k, v in reads.items()}
    # taiyaki/bio.py:43-81: a record with a letter outside the alphabet is dropped (default ACGT;
    # these references hold the modified base Z), or -- filter off -- the letter becomes N
    assert pmf.fasta_file_to_dict(str(fa)) == {k: v[1] for k, v in reads.items() if 'Z' not in v[1]}
    flat = pmf.fasta_file_to_dict(str(fa), filter_ambig=False)
    assert {k: v for k, v in flat.items()} == {k: v[1].replace('Z', 'N') for k, v in reads.items()}
    raw = pmf.fasta_file_to_dict(str(fa), filter_ambig=False, flatten_ambig=False)
    assert raw == {k: v[1] for k, v in reads.items()}
    raw = list(cli.iterate_raw_reads(str(folder)))
    assert [r['read_id'] for r in raw] == ['orphan', 'read0', 'read1', 'read2']
    assert raw[1]['digitisation'] == 8192.0 and raw[1]['dacs'].dtype == np.int16
    assert [r['read_id'] for r in cli.iterate_raw_reads(str(folder), limit=2)] == ['orphan', 'read0']
    ai = cli.make_alphabet_info('ACGT', [['Z', 'C', '5mC'], ['Y', 'A', '6mA']])
    assert (ai.alphabet, ai.collapse_alphabet, ai.mod_long_names) == ('AYCZGT', 'AACCGT', ['6mA', '5mC'])
    assert ai.collapse_sequence('AZYT') == 'ACAT'
    with pytest.raises(AssertionError):
        cli.make_alphabet_info('ACGT', [['C', 'C', 'x']])
    assert pmf.trim_bounds(1000, 100, 20) == (100, 980)
    assert pmf.trim_bounds(100, 80, 30) == (0, 100)          # nothing would be left: no trim
    args = cli.get_parser().parse_args(['--localpen', '1.5', '--max_read_length', 'None', '--mod', 'Z', 'C',
                                        '5mC', 'in', 'p.tsv', 'out.hdf5', 'm.checkpoint', 'r.fa'])
    assert args.localpen == 1.5 and args.max_read_length is None and args.mod == [['Z', 'C', '5mC']]


# ------------------------------------------------------------------ GPU

@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


@pytest.mark.gpu
@pytest.mark.parametrize('tag', list(CASES))
def test_remap_golden(g, dev, tag):
    from taiyaki_b200 import flipflop_remap
    score, path = flipflop_remap.flipflop_remap(g[tag + '_scores'], str(g[tag + '_seq']),
                                                localpen=float(g[tag + '_localpen']))
    assert score == float(g[tag + '_score'])
    np.testing.assert_array_equal(path, g[tag + '_path'])


@pytest.mark.gpu
def test_remap_unit_test_tables(g, dev):
    from taiyaki_b200 import flipflop_remap
    for tag in ('kat1', 'kat2'):
        score, path = flipflop_remap.flipflop_remap(torch.tensor(g[tag + '_scores'], device=dev),
                                                    str(g[tag + '_seq']), alphabet='AB', localpen=-0.5)
        assert score == float(g[tag + '_score'])
        assert path.tolist() == g[tag + '_path'].tolist()
    score2, path2 = flipflop_remap.map_to_crf_viterbi(g['kat1_scores'], [8, 6, 1], [0, 10, 5, 0], localpen=-0.5)
    assert score2 == 6.0 and path2.tolist() == [0, 1, 1, 2, 2, 3, 3]


@pytest.mark.gpu
def test_remap_batch_equals_single(g, dev):
    """All golden reads (ragged T and M) in one launch."""
    from taiyaki_b200 import flipflop_remap
    for pen in (1e30, 0.5):
        tags = [t for t in CASES]
        res = flipflop_remap.flipflop_remap_batch([g[t + '_scores'] for t in tags],
                                                  [str(g[t + '_seq']) for t in tags], localpen=pen)
        from oracle import oracle
        for t, (score, path) in zip(tags, res):
            bases = np.array(['ACGT'.find(b) for b in str(g[t + '_seq'])])
            step, stay = oracle.remap_indices(bases, 4)
            oscore, opath = oracle.map_to_crf_viterbi(g[t + '_scores'], step, stay, pen)
            assert score == oscore
            np.testing.assert_array_equal(path, opath)


@pytest.mark.gpu
def test_remap_six_base_alphabet_short_reference(dev):
    """84 transitions per row with fewer than 64 positions: every score column must be staged
    (round-1 advice: the block had max(64, positions) threads and thread t staged column t)."""
    from oracle import oracle
    from taiyaki_b200 import flipflop_remap
    rng = np.random.RandomState(6)
    alphabet = 'ACGTXY'
    T, L = 120, 30
    seq = ''.join(alphabet[b] for b in rng.randint(0, 6, size=L))
    step, stay = flipflop_remap.remap_indices(seq, alphabet)
    assert max(step.max(), stay.max()) >= 64
    scores = rng.standard_normal((T, 84)).astype('f4')
    score, path = flipflop_remap.flipflop_remap(torch.tensor(scores, device=dev), seq, alphabet=alphabet)
    oscore, opath = oracle.map_to_crf_viterbi(scores, step, stay, 1e30)
    assert score == oscore
    np.testing.assert_array_equal(path, opath)


@pytest.mark.gpu
@pytest.mark.parametrize('T,L,pen', [(8000, 3500, 1e30), (6000, 2500, 2.0), (15000, 13500, 1e30),
                                     (300, 400, 1e30)])
def test_remap_full_size(dev, T, L, pen):
    """Read-sized inputs (the third beyond the shared-memory capacity, i.e. the
    global-memory variant; the last with more positions than blocks, where no complete
    path exists): bit-identical to the oracle, the path is monotone and covers every
    position, and a planted high-scoring global alignment is recovered exactly."""
    from oracle import oracle
    from taiyaki_b200 import flipflop_remap
    rng = np.random.RandomState(T + L)
    bases = rng.randint(0, 4, size=L)
    seq = ''.join('ACGT'[b] for b in bases)
    step, stay = flipflop_remap.remap_indices(seq)
    scores = rng.standard_normal((T, 40)).astype('f4')
    planted = None
    if L < T:      # plant: dwell pattern summing to T, +6 on the planted transitions
        cuts = np.sort(rng.choice(np.arange(1, T), size=L - 1, replace=False))
        pos = np.zeros(T + 1, dtype=int)
        pos[cuts] = 1
        planted = np.cumsum(pos)
        for t in range(T):
            a, b = planted[t], planted[t + 1]
            scores[t, stay[a] if a == b else step[a]] += 6.0
    score, path = flipflop_remap.flipflop_remap(torch.tensor(scores, device=dev), seq, localpen=pen)
    oscore, opath = oracle.map_to_crf_viterbi(scores, step, stay, pen)
    assert score == oscore
    np.testing.assert_array_equal(path, opath)
    if planted is None:
        assert score < -1e29
        return
    p = path[path >= 0]
    assert p[0] == 0 and p[-1] == L - 1 and np.all(np.diff(p) >= 0) and np.all(np.diff(p) <= 1)
    if pen >= 1e29:
        np.testing.assert_array_equal(path, planted)
    else:
        gscore, _ = flipflop_remap.flipflop_remap(scores, seq, localpen=1e30)
        assert score >= gscore


@pytest.mark.gpu
def test_remap_reads_driver(dev):
    """prepare_mapping_funcs.remap_reads on a random-weight network: per read the result
    equals the composition network -> oracle alignment -> from_remapping_path on the same
    transition scores; the reference's failure categories are reported."""
    from oracle import oracle
    from taiyaki_b200 import helpers, prepare_mapping_funcs
    from taiyaki_b200.alphabet import AlphabetInfo
    from taiyaki_b200.mapped_signal_files import check_read
    from taiyaki_b200.prepare_mapping_funcs import RemapResult
    from taiyaki_b200.signal_mapping import SignalMapping
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    torch.manual_seed(4)
    ai = AlphabetInfo('ACGTZ', 'ACGTC', ['5mC'])
    model = helpers.load_model(os.path.join(root, 'models', 'mLstm_flipflop.py'),
                               model_metadata={'reverse': False, 'standardize': True}, stride=5,
                               winlen=19, insize=1, size=64,
                               alphabet_info=AlphabetInfo('ACGT', 'ACGT')).to(dev)
    rng = np.random.RandomState(9)
    reads, params = [], {}
    for i, (n, L) in enumerate(((6000, 500), (3011, 200), (4500, 60))):
        rid = 'read%d' % i
        ref = ''.join('ACGTZ'[b] for b in rng.choice(5, size=L, p=[.25, .15, .25, .25, .1]))
        reads.append({'read_id': rid, 'dacs': rng.randint(300, 700, size=n).astype(np.int16),
                      'offset': 10.0, 'range': 1400.0, 'digitisation': 8192.0, 'ref': ref})
        params[rid] = {'trim_start': 50 * i, 'trim_end': 20, 'shift': 85.0, 'scale': 14.0}
    reads.append({'read_id': 'noref', 'dacs': reads[0]['dacs'], 'offset': 0.0, 'range': 1.0,
                  'digitisation': 1.0, 'ref': None})
    reads.append(dict(reads[0], read_id='noparams'))
    reads.append(dict(reads[0], read_id='long', ref='ACGT' * 300))
    params['long'] = params['read0']
    res = prepare_mapping_funcs.remap_reads(reads, model, params, ai, max_read_length=1000, localpen=0.0)
    assert [r[1] for r in res] == [RemapResult.SUCCESS] * 3 + [
        RemapResult.NO_REF_FOUND, RemapResult.NO_PARAMS, RemapResult.REF_TOO_LONG]
    assert all(r[0] is None for r in res[3:])
    for read, (d, _) in zip(reads[:3], res[:3]):
        p = params[read['read_id']]
        dacs = read['dacs']
        start, end = p['trim_start'], len(dacs) - p['trim_end']
        current = (dacs[start:end] + read['offset']) * read['range'] / read['digitisation']
        sig = ((current - p['shift']) / p['scale']).astype(np.float32)
        with torch.no_grad():
            trans = model(torch.tensor(sig[:, None, None], device=dev))[:, 0].cpu().numpy()
        can = ai.collapse_sequence(read['ref'])
        step, stay = oracle.remap_indices(np.array(['ACGT'.find(b) for b in can]), 4)
        _, opath = oracle.map_to_crf_viterbi(trans, step, stay, 0.0)
        want = SignalMapping.from_remapping_path(
            opath, SignalMapping.get_integer_reference(read['ref'], ai.alphabet), 5, dacs, start)
        np.testing.assert_array_equal(d['Ref_to_signal'], want.Ref_to_signal)
        np.testing.assert_array_equal(d['Reference'], want.Reference)
        assert d['Reference'].max() == 4 and d['read_id'] == read['read_id']
        np.testing.assert_array_equal(d['Dacs'], dacs)
        sm = SignalMapping(**d)
        assert check_read(sm) == 'pass' and sm.scale_frompA == 14.0


@pytest.mark.gpu
def test_prepare_mapped_reads_cli(dev, tmp_path):
    """bin/prepare_mapped_reads.py end to end: raw reads + parameters + references + a
    checkpoint -> a batched mapped-signal file that the reader (and so train_flipflop.py)
    loads; reads without parameters or reference are reported, not written."""
    from taiyaki_b200 import helpers, mapped_signal_files
    from taiyaki_b200.alphabet import AlphabetInfo
    cli = _load_cli('prepare_mapped_reads')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    folder, tsv, fa, reads = _write_remap_inputs(tmp_path)
    torch.manual_seed(4)
    model = helpers.load_model(os.path.join(root, 'models', 'mLstm_flipflop.py'),
                               model_metadata={'reverse': False, 'standardize': True}, stride=5,
                               winlen=19, insize=1, size=64,
                               alphabet_info=AlphabetInfo('ACGT', 'ACGT')).to(dev)
    ckpt, _ = helpers.save_model(model, str(tmp_path))
    out = str(tmp_path / 'mapped.hdf5')
    count, errs = cli.main(['--mod', 'Z', 'C', '5mC', '--reads_per_batch', '2', str(folder), str(tsv),
                            out, ckpt, str(fa)])
    assert count == 3 and sum(errs.values()) == 1
    with mapped_signal_files.MappedSignalReader(out) as msr:
        ai = msr.get_alphabet_information()
        assert (ai.alphabet, ai.collapse_alphabet, ai.mod_long_names) == ('ACZGT', 'ACCGT', ['5mC'])
        assert sorted(msr.get_read_ids()) == ['read0', 'read1', 'read2'] and msr.check() == 'pass'
        for read in msr.reads():
            dacs, ref = reads[read.read_id]
            np.testing.assert_array_equal(read.Dacs, dacs)
            assert ''.join('ACZGT'[b] for b in read.Reference) == ref
            assert read.scale_frompA == 14.0 and read.digitisation == 8192.0
    with pytest.raises(SystemExit):          # refuses to overwrite
        cli.main([str(folder), str(tsv), out, ckpt, str(fa)])
