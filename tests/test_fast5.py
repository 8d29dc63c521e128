"""Raw reads from fast5 files into the data-preparation callers (the input
side of bin/prepare_mapped_reads.py; bin/generate_per_read_params.py; bin/basecall.py).

tests/golden/prepare_remap.npz (make_golden.py prepare) holds the five reads of the
reference's test/data/reads, its per-read parameter table, its references, its shipped
remapping model (bf16-rounded) and the mappings the REFERENCE's own code produces from them
on the CPU in fp32.  The generator pins the decoded samples against a file the reference
wrote from the same fast5 files through ont_fast5_api + h5py.

CPU, build container only (reference tree present): the nine cases of the reference's
test/unit/test_iterate_fast5_reads.py on its own fixture files, samples and parameter table.
CPU, anywhere: the same reads as fast5 files written by tests/fast5_fixture.py.
GPU: bin/prepare_mapped_reads.py and bin/basecall.py from fast5 input with the shipped
remapping model, against the reference's mappings / reference sequences."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fast5_fixture  # noqa: E402

REF_DATA = '/root/reference/test/data'
needs_ref_files = pytest.mark.skipif(not os.path.isdir(REF_DATA),
                                     reason='the reference tree is only in the build container')
EXPECTED_READ_IDS = [
    '0f776a08-1101-41d4-8097-89136494a46e', '1f1a0f33-e2ac-431a-8f48-c3c687a7a7dc',
    'b7096acd-b528-474e-a863-51295d18d3de', 'db6b45aa-5d21-45cf-a435-05fb8f12e839',
    'de1508c4-755b-489e-9ffb-51af35c9a7e6']
MAPPED = [EXPECTED_READ_IDS[0], EXPECTED_READ_IDS[3], EXPECTED_READ_IDS[4]]     # reads with a reference


@pytest.fixture(scope='module')
def g():
    return fast5_fixture.golden()


@pytest.fixture(autouse=True)
def leave_global_rngs_untouched():
    """Building a network draws its initial weights from numpy's and torch's GLOBAL generators
    (layers.py: orthonormal / truncated-normal initialisers, as in the reference) before the
    golden parameters replace them.  Several older tests seed only one of the generators and so
    see whatever state the tests before them left; this file restores both, so that adding or
    removing it does not change what any other test computes."""
    np_state, torch_state = np.random.get_state(), torch.get_rng_state()
    cuda_state = torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None
    yield
    np.random.set_state(np_state)
    torch.set_rng_state(torch_state)
    if cuda_state is not None:
        torch.cuda.set_rng_state_all(cuda_state)


def _load_cli(name):
    import importlib
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    return importlib.import_module(name)


def _ids(found):
    return sorted(rid for _, rid in found)


# ---------------------------------------------------------------- the reference's own files
@needs_ref_files
@pytest.mark.parametrize('folder,strand_list', [
    ('multireads', None), ('reads', None),
    ('multireads', 'basecaller_output/sequencing_summary.txt'),
    ('reads', 'strand_lists/strand_list_single.txt'),
    ('multireads', 'strand_lists/strand_list.txt'),
    ('multireads', 'strand_lists/strand_list_no_filename.txt'),
    ('reads', 'strand_lists/strand_list_no_filename.txt'),
    ('multireads', 'strand_lists/strand_list_no_read_id.txt')])
def test_iterate_fast5_reads_on_reference_fixtures(folder, strand_list):
    """test/unit/test_iterate_fast5_reads.py:32-82, case by case."""
    from taiyaki_b200.fast5utils import iterate_fast5_reads
    sl = None if strand_list is None else os.path.join(REF_DATA, strand_list)
    assert _ids(iterate_fast5_reads(os.path.join(REF_DATA, folder), strand_list=sl)) == EXPECTED_READ_IDS


@needs_ref_files
def test_strand_list_without_header_is_refused():
    """test_iterate_fast5_reads.py:84-92."""
    from taiyaki_b200.fast5utils import iterate_fast5_reads
    with pytest.raises(Exception):
        list(iterate_fast5_reads(os.path.join(REF_DATA, 'multireads'), strand_list=os.path.join(
            REF_DATA, 'strand_lists/invalid_strand_list_no_header.txt')))


@needs_ref_files
@pytest.mark.parametrize('folder', ['reads', 'multireads'])
def test_reference_fast5_files_decode_to_the_golden_reads(g, folder):
    from taiyaki_b200 import fast5utils
    from taiyaki_b200.signal import Signal
    n = 0
    for filename, rid in fast5utils.iterate_fast5_reads(os.path.join(REF_DATA, folder)):
        with fast5utils.get_fast5_file(filename) as f5:
            assert f5.file_type == ('single-read' if folder == 'reads' else 'multi-read')
            read = f5.get_read(rid)
            sig = Signal(read)
            attrs = fast5utils.get_read_attributes(read)
            assert attrs['read_id'].decode() == rid and int(attrs['duration']) == len(sig.untrimmed_dacs)
            if folder == 'reads':       # context tag of the run (multi-read files of this set lack it)
                assert fast5utils.get_filename(read).decode().startswith('minicol615_20190207')
        np.testing.assert_array_equal(sig.untrimmed_dacs, g[rid + '_dacs'])
        assert [sig.offset, sig.range, sig.digitisation, sig.sample_rate] == list(g[rid + '_channel'])
        n += 1
    assert n == 5


@needs_ref_files
def test_generate_per_read_params_reproduces_the_reference_table(tmp_path, capsys):
    """bin/generate_per_read_params.py on test/data/reads gives test/data/readparams.tsv, the
    table the reference's acceptance test feeds to prepare_mapped_reads.py, to the last digit."""
    cli = _load_cli('generate_per_read_params')
    out = tmp_path / 'params.tsv'
    assert cli.main(['--output', str(out), os.path.join(REF_DATA, 'reads')]) == 5
    want = open(os.path.join(REF_DATA, 'readparams.tsv')).read().strip().splitlines()
    got = out.read_text().strip().splitlines()
    assert got[0] == want[0] and sorted(got[1:]) == sorted(want[1:])
    with pytest.raises(SystemExit):                     # existing output is not overwritten
        cli.main(['--output', str(out), os.path.join(REF_DATA, 'reads')])


@needs_ref_files
def test_modified_base_inputs_of_the_reference_acceptance_flow():
    """test/acceptance/test_prepare_remap.py::test_mod_prepare_remap feeds `--mod Z C 5mC --mod Y A 6mA`
    and test/data/per_read_references.mod_bases.fasta: the alphabet built from the flags and the integer
    labels of those references equal what the reference's own alphabet / signal_mapping modules give
    (loaded from the reference tree for this comparison)."""
    import importlib.util

    def reference_module(name):
        spec = importlib.util.spec_from_file_location('reference_' + name, '/root/reference/taiyaki/%s.py' % name)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        return module
    from taiyaki_b200.prepare_mapping_funcs import fasta_file_to_dict
    from taiyaki_b200.signal_mapping import SignalMapping
    cli = _load_cli('prepare_mapped_reads')
    mine = cli.make_alphabet_info('ACGT', [['Z', 'C', '5mC'], ['Y', 'A', '6mA']])
    theirs = reference_module('alphabet').AlphabetInfo('ACGTZY', 'ACGTCA', ['5mC', '6mA'], do_reorder=True)
    assert (mine.alphabet, mine.collapse_alphabet, mine.mod_long_names, str(mine)) == (
        theirs.alphabet, theirs.collapse_alphabet, theirs.mod_long_names, str(theirs))
    assert mine.alphabet == 'AYCZGT' and mine.can_bases == theirs.can_bases == 'ACGT'
    np.testing.assert_array_equal(mine.collapse_labels, theirs.collapse_labels)
    refs = fasta_file_to_dict(os.path.join(REF_DATA, 'per_read_references.mod_bases.fasta'), alphabet=mine.alphabet)
    assert sorted(refs) == sorted(MAPPED)
    ref_sm = reference_module('signal_mapping').SignalMapping
    for rid, seq in refs.items():
        assert 'Z' in seq
        np.testing.assert_array_equal(SignalMapping.get_integer_reference(seq, mine.alphabet),
                                      ref_sm.get_integer_reference(seq, theirs.alphabet))
        assert mine.collapse_sequence(seq) == theirs.collapse_sequence(seq)
    # with the default alphabet these references are dropped, as bio.fasta_file_to_dict drops them
    assert fasta_file_to_dict(os.path.join(REF_DATA, 'per_read_references.mod_bases.fasta')) == {}


# ---------------------------------------------------------------- the same reads, written here
@pytest.mark.parametrize('multi', [False, True])
def test_written_fast5_files_round_trip(g, tmp_path, multi):
    from taiyaki_b200 import fast5utils, maths
    from taiyaki_b200.signal import Signal
    reads_dir, _, _ = fast5_fixture.write_inputs(tmp_path, g, multi)
    found = list(fast5utils.iterate_fast5_reads(reads_dir))
    assert _ids(found) == EXPECTED_READ_IDS
    for filename, rid in found:
        with fast5utils.get_fast5_file(filename) as f5:
            assert f5.file_type == ('multi-read' if multi else 'single-read')
            assert rid in f5.get_read_ids()
            read = f5.get_read(rid)
            sig = Signal(read)
            with pytest.raises(KeyError):
                f5.get_read('not-a-read')
        np.testing.assert_array_equal(sig.untrimmed_dacs, g[rid + '_dacs'])
        assert sig.read_id == rid and sig.untrimmed_dacs.dtype == np.int16
        # bin/generate_per_read_params.py: med_mad of the whole read's current
        assert maths.med_mad(sig.current) == tuple(g[rid + '_params'][2:])
    assert len(list(fast5utils.iterate_fast5_reads(reads_dir, limit=2))) == 2
    one = found[0][0]
    assert _ids(fast5utils.iterate_fast5_reads(one)) == (EXPECTED_READ_IDS if multi else [found[0][1]])


def test_read_loader_keeps_one_file_open(g, tmp_path):
    """Consecutive reads of one multi-read file share one open file (one parse of its root group);
    another file name replaces it."""
    from taiyaki_b200 import fast5utils
    from taiyaki_b200.signal import Signal
    multi_dir, _, _ = fast5_fixture.write_inputs(tmp_path / 'm', g, multi=True)
    single_dir, _, _ = fast5_fixture.write_inputs(tmp_path / 's', g, multi=False)
    batch = os.path.join(multi_dir, 'batch_0.fast5')
    with fast5utils.ReadLoader() as loader:
        a = loader.get_read(batch, EXPECTED_READ_IDS[0])
        opened = loader._file
        b = loader.get_read(batch, EXPECTED_READ_IDS[1])
        assert loader._file is opened
        np.testing.assert_array_equal(Signal(a).untrimmed_dacs, g[EXPECTED_READ_IDS[0] + '_dacs'])
        np.testing.assert_array_equal(Signal(b).untrimmed_dacs, g[EXPECTED_READ_IDS[1] + '_dacs'])
        c = loader.get_read(os.path.join(single_dir, EXPECTED_READ_IDS[2] + '.fast5'), EXPECTED_READ_IDS[2])
        assert loader._file is not opened and Signal(c).read_id == EXPECTED_READ_IDS[2]
        with pytest.raises(KeyError):
            loader.get_read(batch, 'not-a-read')
    assert loader._file is None


def test_strand_lists_on_written_files(g, tmp_path):
    """The three kinds of strand list (fast5utils.py:121-134) and the files they may name."""
    from taiyaki_b200.fast5utils import iterate_fast5_reads
    reads_dir, _, _ = fast5_fixture.write_inputs(tmp_path, g, multi=False)
    a, b = EXPECTED_READ_IDS[0], EXPECTED_READ_IDS[3]

    def strand_list(text):
        p = tmp_path / 'strands.tsv'
        p.write_text(text)
        return str(p)
    only_ids = strand_list('read_id\n{}\n{}\nmissing-read\n'.format(a, b))
    assert _ids(iterate_fast5_reads(reads_dir, strand_list=only_ids)) == [a, b]
    only_files = strand_list('filename\n{}.fast5\nabsent.fast5\n'.format(a))
    assert _ids(iterate_fast5_reads(reads_dir, strand_list=only_files)) == [a]
    pairs = strand_list('filename_fast5\tread_id\n{0}.fast5\t{0}\n{1}.fast5\t{0}\nabsent.fast5\t{1}\n'.format(a, b))
    assert _ids(iterate_fast5_reads(reads_dir, strand_list=pairs)) == [a]     # wrong pairing, missing file
    with pytest.raises(Exception):
        list(iterate_fast5_reads(reads_dir, strand_list=strand_list('{}.fast5\t{}\n'.format(a, a))))


def test_recursive_search_and_unreadable_files(g, tmp_path, capsys):
    from taiyaki_b200.fast5utils import get_fast5_file_list, iterate_fast5_reads
    reads_dir, _, _ = fast5_fixture.write_inputs(tmp_path, g, multi=False)
    sub = os.path.join(reads_dir, 'batch1')
    os.makedirs(sub)
    moved = EXPECTED_READ_IDS[1] + '.fast5'
    os.rename(os.path.join(reads_dir, moved), os.path.join(sub, moved))
    with open(os.path.join(reads_dir, 'broken.fast5'), 'wb') as fh:
        fh.write(b'this is not an HDF5 file')
    assert len(get_fast5_file_list(reads_dir, recursive=False)) == 5        # 4 reads + the broken file
    assert len(get_fast5_file_list(reads_dir, recursive=True)) == 6
    assert _ids(iterate_fast5_reads(reads_dir, recursive=True)) == EXPECTED_READ_IDS
    flat = _ids(iterate_fast5_reads(reads_dir, recursive=False))
    assert flat == [r for r in EXPECTED_READ_IDS if r != EXPECTED_READ_IDS[1]]
    assert 'skipped this read' in capsys.readouterr().err                   # the broken file is reported


def test_signal_trimming_and_units(g):
    """taiyaki/signal.py:77-123."""
    from taiyaki_b200.signal import Signal
    rid = EXPECTED_READ_IDS[0]
    dacs = g[rid + '_dacs']
    offset, rng, digitisation, rate = g[rid + '_channel']
    info = {'offset': offset, 'range': rng, 'digitisation': digitisation, 'sampling_rate': rate}
    t0, t1, shift, scale = g[rid + '_params']
    params = {'trim_start': int(t0), 'trim_end': int(t1), 'shift': shift, 'scale': scale}
    sig = Signal(dacs=dacs, channel_info=info, read_id=rid, read_params=params)
    assert (sig.signalstart, sig.signalend_exc) == (200, len(dacs) - 50)
    np.testing.assert_array_equal(sig.dacs, dacs[200:-50])
    np.testing.assert_array_equal(sig.untrimmed_current, (dacs + offset) * rng / digitisation)
    np.testing.assert_array_equal(sig.current, sig.untrimmed_current[200:-50])
    np.testing.assert_array_equal(sig.standardized_current, (sig.current - shift) / scale)
    sig.dacs[0] = 0                                     # a copy: the stored samples are untouched
    assert sig.untrimmed_dacs[200] == dacs[200]
    short = Signal(dacs=dacs[:250], channel_info=info, read_params=params)   # nothing would be left
    assert (short.signalstart, short.signalend_exc) == (0, 250)
    with pytest.raises(Exception):
        Signal(dacs=dacs, channel_info=info, read_params=dict(params, trim_end=-1))
    with pytest.raises(Exception):
        Signal()
    plain = Signal(dacs=np.arange(10))
    np.testing.assert_array_equal(plain.standardized_current, np.arange(10))


def test_generate_per_read_params_cli(g, tmp_path, capsys):
    cli = _load_cli('generate_per_read_params')
    reads_dir, tsv, _ = fast5_fixture.write_inputs(tmp_path, g, multi=True)
    out = tmp_path / 'out.tsv'
    assert cli.main(['--output', str(out), '--trim', '200', '50', reads_dir]) == 5
    assert sorted(out.read_text().splitlines()) == sorted(open(tsv).read().splitlines())
    assert cli.main(['--limit', '2', '--trim', '10', '0', reads_dir]) == 2
    rows = capsys.readouterr().out.strip().splitlines()
    assert rows[0].split('\t') == ['UUID', 'trim_start', 'trim_end', 'shift', 'scale']
    assert [r.split('\t')[1:3] for r in rows[1:]] == [['10', '0']] * 2


def test_prepare_cli_reads_fast5_input(g, tmp_path):
    """bin/prepare_mapped_reads.py's read iteration on fast5 input: the dictionaries remap_reads
    takes; samples of reads that will be rejected anyway are not loaded."""
    cli = _load_cli('prepare_mapped_reads')
    reads_dir, _, _ = fast5_fixture.write_inputs(tmp_path, g, multi=False)
    raw = {r['read_id']: r for r in cli.iterate_raw_reads(reads_dir)}
    assert sorted(raw) == EXPECTED_READ_IDS
    for rid, r in raw.items():
        np.testing.assert_array_equal(r['dacs'], g[rid + '_dacs'])
        assert [r['offset'], r['range'], r['digitisation']] == list(g[rid + '_channel'][:3])
    some = list(cli.iterate_raw_reads(reads_dir, wanted=lambda rid: rid in MAPPED))
    assert sorted(r['read_id'] for r in some if r['dacs'] is not None) == MAPPED
    assert len(some) == 5 and len(list(cli.iterate_raw_reads(reads_dir, limit=3))) == 3


def test_basecall_cli_reads_fast5_input(g, tmp_path):
    """bin/basecall.py's signal iteration on fast5 input: the read's current in pA
    (bin/basecall.py:92-116), None for a read that cannot be loaded."""
    cli = _load_cli('basecall')
    reads_dir, _, _ = fast5_fixture.write_inputs(tmp_path, g, multi=True)
    got = dict(cli.iterate_signals(reads_dir))
    assert sorted(got) == EXPECTED_READ_IDS
    for rid, current in got.items():
        offset, rng, digitisation, _ = g[rid + '_channel']
        np.testing.assert_array_equal(current, (g[rid + '_dacs'] + offset) * rng / digitisation)
    assert cli.get_signal(os.path.join(reads_dir, 'batch_0.fast5'), 'not-a-read') is None


def test_prepare_cli_shards_reads_by_position(g, tmp_path):
    """--shard index count / torchrun ranks: disjoint shares that together are the input order;
    the samples of other shards' reads are not touched."""
    import argparse
    cli = _load_cli('prepare_mapped_reads')
    reads_dir, _, _ = fast5_fixture.write_inputs(tmp_path, g, multi=True)
    order = [r['read_id'] for r in cli.iterate_raw_reads(reads_dir)]
    shares = [[r['read_id'] for r in cli.iterate_raw_reads(reads_dir, shard=(i, 3))] for i in range(3)]
    assert shares == [order[0::3], order[1::3], order[2::3]]
    assert [r['read_id'] for r in cli.iterate_raw_reads(reads_dir, limit=3, shard=(1, 2))] == order[1:3:2]
    ns = argparse.Namespace
    assert cli.shard_of_process(ns(shard=None), {}) == (None, False)
    assert cli.shard_of_process(ns(shard=[1, 4]), {'WORLD_SIZE': '8', 'RANK': '5'}) == ((1, 4), False)
    assert cli.shard_of_process(ns(shard=[0, 1]), {}) == (None, False)
    assert cli.shard_of_process(ns(shard=None), {'WORLD_SIZE': '8', 'RANK': '5'}) == ((5, 8), True)
    assert cli.shard_of_process(ns(shard=None), {'WORLD_SIZE': '1', 'RANK': '0'}) == (None, False)
    with pytest.raises(SystemExit):
        cli.shard_of_process(ns(shard=[2, 2]), {})
    assert cli.shard_filename('out.hdf5', (2, 8)) == 'out.hdf5.shard2of8'


def _prepare_rank(rank, world, port, workdir, failed):
    """One torchrun-style rank of bin/prepare_mapped_reads.py on the CPU: the device work
    (network + alignment) is replaced by a stand-in mapping, everything else -- sharding of the
    input, per-shard file, barrier, join on rank 0 -- is the script's own."""
    try:
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                          LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
        import taiyaki_b200.helpers as helpers
        from taiyaki_b200.prepare_mapping_funcs import RemapResult
        from taiyaki_b200.signal_mapping import SignalMapping
        cli = _load_cli('prepare_mapped_reads')

        def fake_remap(reads, model, params, alphabet_info, max_read_length, localpen, stride):
            out = []
            for read in reads:
                if read.get('ref') is None:
                    out.append((None, RemapResult.NO_REF_FOUND))
                    continue
                labels = SignalMapping.get_integer_reference(read['ref'], alphabet_info.alphabet)
                bounds = np.linspace(0, len(read['dacs']), len(labels) + 1).astype(np.int32)
                p = params[read['read_id']]
                out.append((SignalMapping(read['dacs'], bounds, labels, read_id=read['read_id'],
                                          shift_frompA=p['shift'], scale_frompA=p['scale'], range=read['range'],
                                          offset=read['offset'], digitisation=read['digitisation']
                                          ).get_read_dictionary(), RemapResult.SUCCESS))
            return out
        cli.remap_reads = fake_remap
        helpers.load_model = lambda *a, **k: type('NoModel', (), {'to': lambda self, device: self})()
        helpers.guess_model_stride = lambda model: 4
        torch.cuda.set_device = lambda device: None
        cli.main([os.path.join(workdir, 'reads'), os.path.join(workdir, 'readparams.tsv'),
                  os.path.join(workdir, 'mapped.hdf5'), 'unused.checkpoint', os.path.join(workdir, 'refs.fasta')])
    except BaseException as e:                                     # noqa: B902 -- reported to the parent
        failed[rank] = repr(e)
        raise


def test_prepare_cli_under_torchrun_gloo_world2(g, tmp_path):
    """Two ranks: each remaps its share into <output>.shard<r>of2, rank 0 joins them after the
    barrier and removes the shard files."""
    import torch.multiprocessing as mp
    from taiyaki_b200 import mapped_signal_files
    fast5_fixture.write_inputs(tmp_path, g, multi=False)
    ctx = mp.get_context('spawn')
    failed = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_prepare_rank, args=(r, 2, port, str(tmp_path), failed)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert [p.exitcode for p in procs] == [0, 0], dict(failed)
    assert sorted(os.listdir(tmp_path)) == ['mapped.hdf5', 'readparams.tsv', 'reads', 'refs.fasta']
    with mapped_signal_files.MappedSignalReader(str(tmp_path / 'mapped.hdf5')) as msr:
        assert sorted(msr.get_read_ids()) == MAPPED and msr.check() == 'pass'
        for read in msr.reads():
            np.testing.assert_array_equal(read.Dacs, g[read.read_id + '_dacs'])
            np.testing.assert_array_equal(read.Reference, g[read.read_id + '_Reference'])


def test_remap_reads_reports_unloadable_reads():
    """A read whose samples could not be loaded is READ_ID_INFO_NOT_FOUND
    (prepare_mapping_funcs.py:62-68) -- after the checks that need no samples, as in the
    reference.  No device work is reached."""
    from taiyaki_b200 import prepare_mapping_funcs as pmf
    from taiyaki_b200.alphabet import AlphabetInfo

    class NoModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
    reads = [{'read_id': 'a', 'dacs': None, 'ref': 'ACGT'}, {'read_id': 'b', 'dacs': None, 'ref': None},
             {'read_id': 'c', 'dacs': None, 'ref': 'ACGT'}]
    params = {'a': {'trim_start': 0, 'trim_end': 0, 'shift': 0.0, 'scale': 1.0}}
    res = pmf.remap_reads(reads, NoModel(), params, AlphabetInfo('ACGT', 'ACGT'), model_stride=4)
    assert [r[1] for r in res] == [pmf.RemapResult.READ_ID_INFO_NOT_FOUND, pmf.RemapResult.NO_REF_FOUND,
                                   pmf.RemapResult.NO_PARAMS]


# ---------------------------------------------------------------- GPU: the flows end to end
@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


def remapping_model(g, dev):
    """The reference's shipped mGru_flipflop remapping model (size 96, stride 4) from the golden
    file's bf16-rounded parameters."""
    from taiyaki_b200 import helpers
    from taiyaki_b200.alphabet import AlphabetInfo
    size, stride, winlen = (int(v) for v in g['model_cfg'])
    model = helpers.load_model(os.path.join(ROOT, 'models', 'mGru_flipflop.py'),
                               model_metadata={'reverse': False, 'standardize': True}, size=size,
                               stride=stride, winlen=winlen, insize=1,
                               alphabet_info=AlphabetInfo('ACGT', 'ACGT'))
    state = {k[len('param_'):]: torch.from_numpy(g[k].view(np.int16).copy()).view(torch.bfloat16).float()
             for k in g.files if k.startswith('param_')}
    model.load_state_dict(state)
    return model.to(dev)


def _mapping_agreement(got, want, stride):
    d = np.abs(np.asarray(got, dtype=np.int64) - np.asarray(want, dtype=np.int64))
    return float((d == 0).mean()), float((d <= stride).mean()), int(d.max())


@pytest.mark.gpu
@pytest.mark.parametrize('precision,multi', [('fp32', False), ('bf16', True)])
def test_prepare_mapped_reads_from_fast5_matches_reference_flow(g, dev, tmp_path, precision, multi):
    """The reference's acceptance flow (test/acceptance/test_prepare_remap.py: reads + parameter
    table + references + remapping model -> mapped-signal file) from fast5 input, against the
    mappings the reference's own code computes from the same inputs on the CPU in fp32
    (make_golden.py prepare).  Samples, labels and scalars are exact.  Ref_to_signal goes through
    the network: in the fp32 parity mode the scores differ from the CPU's by summation order only
    and every boundary of the three reads is the same (B200, profiles/r2_fast5_prepare_gpu.log);
    with bf16 products 0 to 0.7 % of the 2100 - 3100 boundaries per read move, by at most 24
    samples (stride 4) -- alignment is a max over paths, near-ties flip."""
    from taiyaki_b200 import helpers, layers, mapped_signal_files
    cli = _load_cli('prepare_mapped_reads')
    reads_dir, tsv, fasta = fast5_fixture.write_inputs(tmp_path, g, multi)
    ckpt, _ = helpers.save_model(remapping_model(g, dev), str(tmp_path))
    out = str(tmp_path / 'mapped.hdf5')
    stride = int(g['model_cfg'][1])
    layers.set_precision(precision)
    try:
        count, errs = cli.main(['--reads_per_batch', '2', reads_dir, tsv, out, ckpt, fasta])
    finally:
        layers.set_precision('bf16')
    assert count == 3 and sum(errs.values()) == 2               # two reads have no reference
    with mapped_signal_files.MappedSignalReader(out) as msr:
        assert msr.check() == 'pass' and sorted(msr.get_read_ids()) == MAPPED
        for read in msr.reads():
            rid = read.read_id
            np.testing.assert_array_equal(read.Dacs, g[rid + '_dacs'])
            np.testing.assert_array_equal(read.Reference, g[rid + '_Reference'])
            t0, t1, shift, scale = g[rid + '_params']
            offset, rng, digitisation, _ = g[rid + '_channel']
            assert [read.shift_frompA, read.scale_frompA, read.range, read.offset, read.digitisation] == [
                shift, scale, rng, offset, digitisation]
            same, near, worst = _mapping_agreement(read.Ref_to_signal, g[rid + '_Ref_to_signal'], stride)
            print('%s %s: identical %.4f  within one block %.4f  max %d samples' % (precision, rid[:8], same, near, worst))
            if precision == 'fp32':
                assert same >= 0.999 and near >= 0.999, (same, near, worst)     # measured: 1.0000 on all three
            else:
                assert same >= 0.98 and near >= 0.99, (same, near, worst)       # measured: 0.9931 .. 1.0000
            # what the reference's acceptance test checks: a chunk with a plausible dwell
            chunk = read.get_chunk_with_sample_length(1000, start_sample=10000)
            assert 7 < chunk.sig_len / (chunk.seq_len + 0.0001) < 13


@pytest.mark.gpu
def test_prepare_mapped_reads_shards_join_to_the_unsharded_result(g, dev, tmp_path):
    """--shard 0 2 and --shard 1 2 joined by misc/merge_mappedsignalfiles.py hold the reads of one
    unsharded run, with the same mappings (each read is remapped on its own)."""
    from taiyaki_b200 import helpers, mapped_signal_files
    cli = _load_cli('prepare_mapped_reads')
    sys.path.insert(0, os.path.join(ROOT, 'misc'))
    import importlib
    merge = importlib.import_module('merge_mappedsignalfiles')
    reads_dir, tsv, fasta = fast5_fixture.write_inputs(tmp_path, g, multi=False)
    ckpt, _ = helpers.save_model(remapping_model(g, dev), str(tmp_path))
    whole, out = str(tmp_path / 'whole.hdf5'), str(tmp_path / 'sharded.hdf5')
    assert cli.main([reads_dir, tsv, whole, ckpt, fasta])[0] == 3
    counts = [cli.main(['--shard', str(i), '2', reads_dir, tsv, out, ckpt, fasta])[0] for i in range(2)]
    assert sum(counts) == 3 and min(counts) >= 1
    shards = [cli.shard_filename(out, (i, 2)) for i in range(2)]
    assert merge.main([out, '--input', shards[0], 'None', '--input', shards[1], 'None']) == 3
    with mapped_signal_files.MappedSignalReader(whole) as a, mapped_signal_files.MappedSignalReader(out) as b:
        assert sorted(a.get_read_ids()) == sorted(b.get_read_ids()) == MAPPED
        for rid in MAPPED:
            np.testing.assert_array_equal(a.get_read(rid).Ref_to_signal, b.get_read(rid).Ref_to_signal)
            np.testing.assert_array_equal(a.get_read(rid).Dacs, b.get_read(rid).Dacs)


@pytest.mark.gpu
def test_basecall_from_fast5_recovers_the_reference_sequences(g, dev, tmp_path):
    """bin/basecall.py on fast5 input with the shipped remapping model and the reference's scaling
    table: the calls of real reads agree with the reads' known reference sequences to the extent
    a small r9 model does (well above 80 % of the reference found in the call)."""
    import difflib
    from taiyaki_b200 import helpers
    cli = _load_cli('basecall')
    reads_dir, tsv, _ = fast5_fixture.write_inputs(tmp_path, g, multi=True)
    ckpt, _ = helpers.save_model(remapping_model(g, dev), str(tmp_path))
    out = tmp_path / 'calls.fa'
    cli.main(['--scaling', tsv, '--output', str(out), reads_dir, ckpt])
    lines = out.read_text().strip().splitlines()
    calls = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines), 2)}
    assert sorted(calls) == EXPECTED_READ_IDS
    for rid in MAPPED:
        ref, call = str(g[rid + '_reference']), calls[rid]
        # the reference sequence may cover only part of the read (read 0f776a08 maps from sample
        # 7917 of 28005), so the measure is the share of the REFERENCE found in the call
        assert set(call) <= set('ACGT') and 0.8 * len(ref) < len(call) < 2.0 * len(ref), (len(ref), len(call))
        blocks = difflib.SequenceMatcher(None, ref, call, autojunk=False).get_matching_blocks()
        found = sum(b.size for b in blocks) / len(ref)
        print('%s: reference %d bases, call %d bases, %.3f of the reference found in the call' % (
            rid[:8], len(ref), len(call), found))
        assert found > 0.8, found
