"""misc/merge_mappedsignalfiles.py (the file format's tooling): the cases of
the reference's test/acceptance/test_merge_mappedsignalfiles.py (usage, merging its two fixture
files -- build container only) and the alphabet rules of misc/merge_mappedsignalfiles.py:64-168
on files written here from the golden real reads."""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DATA = '/root/reference/test/data/mapped_signal_file'
needs_ref_files = pytest.mark.skipif(not os.path.isdir(REF_DATA),
                                     reason='the reference tree is only in the build container')


@pytest.fixture(scope='module')
def cli():
    sys.path.insert(0, os.path.join(ROOT, 'misc'))
    return importlib.import_module('merge_mappedsignalfiles')


@pytest.fixture(scope='module')
def reads():
    sys.path.insert(0, HERE)
    from test_real_reads import golden_reads
    return golden_reads(np.load(os.path.join(HERE, 'golden', 'real_reads.npz')))


def write(path, reads, info):
    from taiyaki_b200.mapped_signal_files import MappedSignalWriter
    with MappedSignalWriter(str(path), info) as msw:
        for r in reads:
            msw.write_read(r if isinstance(r, dict) else r.get_read_dictionary())
    return str(path)


def load(path):
    from taiyaki_b200.mapped_signal_files import MappedSignalReader
    with MappedSignalReader(str(path)) as msr:
        assert msr.check() == 'pass'
        return msr.get_alphabet_information(), {r.read_id: r for r in msr.reads()}


def test_usage(cli, capsys):
    """test_merge_mappedsignalfiles.py:36-40."""
    with pytest.raises(SystemExit) as e:
        cli.main([])
    assert e.value.code == 2 and 'usage' in capsys.readouterr().err


@needs_ref_files
def test_merge_reference_fixture_files(cli, tmp_path, reads):
    """test_merge_mappedsignalfiles.py:60-137: two per-read files in, every read out."""
    out = tmp_path / 'merged.hdf5'
    n = cli.main([str(out), '--input', os.path.join(REF_DATA, 'mapped_reads_0.hdf5'), 'None',
                  '--input', os.path.join(REF_DATA, 'mapped_reads_1.hdf5'), 'None', '--batch_format'])
    info, merged = load(out)
    assert n == 7 and len(merged) == 7 and info.alphabet == 'ACGT'
    for r in reads:
        m = merged[r.read_id]
        np.testing.assert_array_equal(m.Dacs, r.Dacs)
        np.testing.assert_array_equal(m.Ref_to_signal, r.Ref_to_signal)
        np.testing.assert_array_equal(m.Reference, r.Reference)
        assert (m.shift_frompA, m.scale_frompA, m.range, m.offset, m.digitisation) == (
            r.shift_frompA, r.scale_frompA, r.range, r.offset, r.digitisation)


def test_merge_batched_shards_duplicates_and_limits(cli, tmp_path, reads, capsys):
    from taiyaki_b200.alphabet import AlphabetInfo
    acgt = AlphabetInfo('ACGT', 'ACGT')
    a = write(tmp_path / 'a.hdf5', reads[:3], acgt)
    b = write(tmp_path / 'b.hdf5', reads[2:], acgt)            # read 2 is in both shards
    out = tmp_path / 'all.hdf5'
    assert cli.main([str(out), '--input', a, 'None', '--input', b, 'None']) == len(reads)
    assert '1 reads found in previous file' in capsys.readouterr().err
    _, merged = load(out)
    assert sorted(merged) == sorted(r.read_id for r in reads)
    np.testing.assert_array_equal(merged[reads[2].read_id].Dacs, reads[2].Dacs)
    # limits: that many reads of each input, the choice repeatable under --seed
    picks = []
    for k in range(2):
        lim = tmp_path / ('lim%d.hdf5' % k)
        assert cli.main([str(lim), '--seed', '5', '--input', a, '2', '--input', b, '1']) == 3
        picks.append(sorted(load(lim)[1]))
    assert picks[0] == picks[1] and len(set(picks[0])) == 3


def test_alphabets_must_agree_without_mod_merge(cli, tmp_path, reads, capsys):
    from taiyaki_b200.alphabet import AlphabetInfo
    a = write(tmp_path / 'a.hdf5', reads[:2], AlphabetInfo('ACGT', 'ACGT'))
    b = write(tmp_path / 'b.hdf5', reads[2:4], AlphabetInfo('ACGTZ', 'ACGTC', ['5mC']))
    with pytest.raises(SystemExit) as e:
        cli.main([str(tmp_path / 'o.hdf5'), '--input', a, 'None', '--input', b, 'None'])
    assert e.value.code == 1 and 'differs from that in' in capsys.readouterr().err
    assert not os.path.exists(tmp_path / 'o.hdf5')            # refused before anything is written


def _with_labels(read, mapping):
    d = read.get_read_dictionary()
    d['Reference'] = np.array([mapping.get(int(x), int(x)) for x in d['Reference']], dtype=np.int16)
    return d


def test_mod_merge_recodes_labels(cli, tmp_path, reads):
    """--allow_mod_merge (merge_mappedsignalfiles.py:64-131,156-168): ACGT + 5mC and ACGT + 6mA
    give one alphabet with both, sorted as AlphabetInfo(do_reorder=True) sorts, and every label
    still names the base it named in its own file."""
    from taiyaki_b200.alphabet import AlphabetInfo
    mc = AlphabetInfo('ACGTZ', 'ACGTC', ['5mC'])
    ma = AlphabetInfo('ACGTY', 'ACGTA', ['6mA'])
    a = write(tmp_path / 'a.hdf5', [_with_labels(reads[0], {1: 4})], mc)       # every C -> Z
    b = write(tmp_path / 'b.hdf5', [_with_labels(reads[1], {0: 4})], ma)       # every A -> Y
    c = write(tmp_path / 'c.hdf5', [reads[2]], AlphabetInfo('ACGT', 'ACGT'))
    out = tmp_path / 'merged.hdf5'
    assert cli.main([str(out), '--allow_mod_merge', '--input', a, 'None', '--input', b, 'None',
                     '--input', c, 'None']) == 3
    info, merged = load(out)
    assert (info.alphabet, info.collapse_alphabet, info.mod_long_names) == ('AYCZGT', 'AACCGT', ['6mA', '5mC'])
    acgt = AlphabetInfo('ACGT', 'ACGT')
    for src, src_info, relabel in ((reads[0], mc, {1: 4}), (reads[1], ma, {0: 4}), (reads[2], acgt, {})):
        want = ''.join(src_info.alphabet[relabel.get(int(x), int(x))] for x in src.Reference)
        assert ''.join(info.alphabet[x] for x in merged[src.read_id].Reference) == want
    assert 'Z' in ''.join(info.alphabet[x] for x in merged[reads[0].read_id].Reference)


@pytest.mark.parametrize('second,message', [
    (('ACGTZ', 'ACGTA', ['5mC']), 'Incompatible modified bases'),        # Z under another canonical base
    (('ACGTZ', 'ACGTC', ['5hmC']), 'Incompatible modified bases'),       # Z with another long name
    (('ACGTY', 'ACGTC', ['5mC']), 'Incompatible modified bases'),        # the long name under another letter
    (('ACGU', 'ACGU', []), 'All canonical alphabets must be the same')])
def test_incompatible_alphabets_are_refused(cli, tmp_path, reads, capsys, second, message):
    from taiyaki_b200.alphabet import AlphabetInfo
    a = write(tmp_path / 'a.hdf5', reads[:1], AlphabetInfo('ACGTZ', 'ACGTC', ['5mC']))
    b = write(tmp_path / 'b.hdf5', reads[1:2], AlphabetInfo(*second))
    with pytest.raises(SystemExit) as e:
        cli.main([str(tmp_path / 'o.hdf5'), '--allow_mod_merge', '--input', a, 'None', '--input', b, 'None'])
    assert e.value.code == 1 and message in capsys.readouterr().err
