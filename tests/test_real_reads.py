"""Real r9.4.1 mapped reads (SURVEY 8(c) "real-data fixtures", 8(f) row 4).

tests/golden/real_reads.npz (make_golden.py real) holds the 7 reads of the reference's
test/data/mapped_signal_file/mapped_reads_{0,1}.hdf5 -- each accepted by the reference's own
SignalMapping.check() -- together with the reference's chunk sampling on them under fixed
numpy seeds and the reference C loss on their real label sequences.

CPU: the mapped-signal readers over the plain-Python HDF5 decoder (only where the
reference's files are present, i.e. in the build container), the host chunk sampling
bit-identical to the reference's.  GPU: device batch assembly on the real reads against
the host path, the CRF kernels on real label sequences against the reference C."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden', 'real_reads.npz')
REF_DATA = '/root/reference/test/data/mapped_signal_file'
needs_ref_files = pytest.mark.skipif(not os.path.isdir(REF_DATA),
                                     reason='the reference tree is only in the build container')


@pytest.fixture(scope='module')
def g():
    return np.load(GOLDEN)


def golden_reads(g):
    from taiyaki_b200.signal_mapping import SignalMapping
    reads = []
    for rid in g['read_ids']:
        rid = str(rid)
        shift, scale, rng, offset, digitisation = g['read_%s_attrs' % rid]
        reads.append(SignalMapping(g['read_%s_Dacs' % rid], g['read_%s_Ref_to_signal' % rid],
                                   g['read_%s_Reference' % rid], read_id=rid, shift_frompA=shift,
                                   scale_frompA=scale, range=rng, offset=offset,
                                   digitisation=digitisation))
    return reads


@needs_ref_files
def test_mapped_signal_reader_on_reference_files(g):
    from taiyaki_b200 import mapped_signal_files
    n = 0
    for fn in ('mapped_reads_0', 'mapped_reads_1'):
        with mapped_signal_files.MappedSignalReader(os.path.join(REF_DATA, fn + '.hdf5')) as msr:
            assert isinstance(msr, mapped_signal_files.PerReadHDF5Reader)
            assert msr.version == 8
            ai = msr.get_alphabet_information()
            assert ai.alphabet == 'ACGT' and ai.collapse_alphabet == 'ACGT' and ai.nbase == 4
            ids = msr.get_read_ids()
            assert ids == [str(x) for x in g[fn + '_read_ids']]
            assert msr.check() == 'pass'
            for read in msr.reads():
                rid = read.read_id
                np.testing.assert_array_equal(read.Dacs, g['read_%s_Dacs' % rid])
                np.testing.assert_array_equal(read.Ref_to_signal, g['read_%s_Ref_to_signal' % rid])
                np.testing.assert_array_equal(read.Reference, g['read_%s_Reference' % rid])
                assert [read.shift_frompA, read.scale_frompA, read.range, read.offset,
                        read.digitisation] == list(g['read_%s_attrs' % rid])
                n += 1
            some = list(msr.reads(ids[-1:] + ['not-a-read']))
            assert [r.read_id for r in some] == ids[-1:]
            assert msr.get_read(ids[0]).read_id == ids[0]
    assert n == 7
    # the third fixture: attributes stored in another order, trimmed mapping
    with mapped_signal_files.HDF5Reader(os.path.join(REF_DATA, 'mapped_remap_samref.hdf5')) as msr:
        assert len(msr.get_read_ids()) == 3 and msr.check() == 'pass'


@needs_ref_files
def test_train_entry_point_loads_hdf5(tmp_path):
    """bin/train_flipflop.py load_data on a mapped-signal file, with --limit and a strand list."""
    import argparse
    import importlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'bin'))
    tf = importlib.import_module('train_flipflop')
    log = type('L', (), {'write': lambda self, m: None})()
    strands = tmp_path / 'strands.tsv'
    strands.write_text('filename\tread_id\nx.fast5\t34bebf28-d997-446e-8dda-ce707a266c2d\n'
                       'y.fast5\tcb8a6688-0ddb-42ff-ad7e-de004d738790\n')
    args = argparse.Namespace(input=os.path.join(REF_DATA, 'mapped_reads_1.hdf5'), limit=None,
                              input_strand_list=None, mod_factor=[8.0, 1.0, 50000])
    reads, ai, mod_info = tf.load_data(args, log, None)
    assert len(reads) == 5 and ai.alphabet == 'ACGT' and list(mod_info.mod_cat_weights) == [1.0] * 4
    args.limit = 2
    assert len(tf.load_data(args, log, None)[0]) == 2
    args.limit, args.input_strand_list = None, str(strands)
    assert sorted(r.read_id for r in tf.load_data(args, log, None)[0]) == [
        '34bebf28-d997-446e-8dda-ce707a266c2d', 'cb8a6688-0ddb-42ff-ad7e-de004d738790']


def test_batched_mapped_signal_file(g, tmp_path):
    """The batched layout (the reference writer's default: concatenated arrays per batch,
    chunked + shuffle + deflate, variable-length read ids) written by the minimal writer
    from the real reads and read back through BatchHDF5Reader."""
    from taiyaki_b200 import mapped_signal_files
    from taiyaki_b200.hdf5_min_write import write_batched_mapped_signal_file
    reads = golden_reads(g)
    fn = str(tmp_path / 'batched.hdf5')
    write_batched_mapped_signal_file(fn, reads, batch_size=3, chunk=7000)
    # the same with 500-element chunks: several hundred chunks per dataset, i.e. chunk B-trees of
    # two levels (64 entries per node)
    fn_small = str(tmp_path / 'batched_small_chunks.hdf5')
    write_batched_mapped_signal_file(fn_small, reads, batch_size=4, chunk=500)
    with mapped_signal_files.MappedSignalReader(fn_small) as msr:
        for a, b in zip(msr.reads(), reads):
            np.testing.assert_array_equal(a.Dacs, b.Dacs)
            np.testing.assert_array_equal(a.Ref_to_signal, b.Ref_to_signal)
    with mapped_signal_files.MappedSignalReader(fn) as msr:
        assert isinstance(msr, mapped_signal_files.BatchHDF5Reader)
        assert msr.batch_names == ['Batch_0', 'Batch_1', 'Batch_2']
        assert msr.get_read_ids() == [r.read_id for r in reads]
        assert msr.get_alphabet_information().alphabet == 'ACGT' and msr.check() == 'pass'
        got = list(msr.reads())
        assert [r.read_id for r in got] == [r.read_id for r in reads]
        for a, b in zip(got, reads):
            np.testing.assert_array_equal(a.Dacs, b.Dacs)
            np.testing.assert_array_equal(a.Ref_to_signal, b.Ref_to_signal)
            np.testing.assert_array_equal(a.Reference, b.Reference)
            assert (a.shift_frompA, a.scale_frompA, a.range, a.offset, a.digitisation) == (
                b.shift_frompA, b.scale_frompA, b.range, b.offset, b.digitisation)
        pick = [reads[5].read_id, reads[0].read_id]
        assert sorted(r.read_id for r in msr.reads(pick)) == sorted(pick)
        np.testing.assert_array_equal(msr.get_read(reads[4].read_id).Dacs, reads[4].Dacs)


def test_batched_file_with_tens_of_thousands_of_reads(tmp_path):
    """Read ids are variable-length strings in global heap collections: the writer adds collections
    as they fill up (round-1 advice: a single 1 MiB collection overflowed at ~9.3k reads, at close(),
    after all the remapping work) and stores an id once although two datasets reference it."""
    from taiyaki_b200 import mapped_signal_files
    from taiyaki_b200.hdf5_min_write import Writer, write_batched_mapped_signal_file
    n = 40000
    rng = np.random.RandomState(0)
    reads = [dict(read_id='%08x-%04x-4%03x-a%03x-%012x' % tuple(rng.randint(0, 2 ** 31, size=5) % (
                      16 ** 8, 16 ** 4, 16 ** 3, 16 ** 3, 16 ** 12)),
                  Dacs=np.arange(6, dtype=np.int16) + i % 7, Ref_to_signal=np.array([0, 2, 4, 6], dtype=np.int32),
                  Reference=np.array([0, 1, 2], dtype=np.int16), shift_frompA=0.0, scale_frompA=1.0,
                  range=1.0, offset=0.0, digitisation=1.0) for i in range(n)]
    fn = str(tmp_path / 'many.hdf5')
    write_batched_mapped_signal_file(fn, reads, batch_size=1000, chunk=4096)
    with mapped_signal_files.MappedSignalReader(fn) as msr:
        ids = msr.get_read_ids()
        # (batch groups are name-ordered, Batch_10 before Batch_2, in the reference's files too)
        assert len(ids) == n and sorted(ids) == sorted(r['read_id'] for r in reads)
        assert len(msr.batch_names) == 40
        last = msr.get_read(reads[-1]['read_id'])
        np.testing.assert_array_equal(last.Dacs, reads[-1]['Dacs'])
        got = [r.read_id for r in msr.reads([reads[12345]['read_id'], reads[777]['read_id']])]
        assert sorted(got) == sorted([reads[12345]['read_id'], reads[777]['read_id']])
    w = Writer()
    w.vlen_refs(['x' * 40 + str(i) for i in range(70000)] * 2)
    assert len(w.collections) >= 2 and all(len(c[1]) <= 65535 for c in w.collections)
    assert sum(len(c[1]) for c in w.collections) == 70000          # stored once, referenced twice


def test_mapped_signal_writer_round_trip(g, tmp_path):
    """MappedSignalWriter (read dictionaries in, batched file out) -> MappedSignalReader,
    with a modified-base alphabet and one batch per two reads."""
    from taiyaki_b200 import mapped_signal_files
    from taiyaki_b200.alphabet import AlphabetInfo
    reads = golden_reads(g)
    fn = str(tmp_path / 'written.hdf5')
    ai = AlphabetInfo('ACGTZY', 'ACGTCA', ['5mC', '6mA'])
    with mapped_signal_files.MappedSignalWriter(fn, ai) as msw:
        msw.batch_size = 2
        for r in reads:
            msw.write_read(r.get_read_dictionary())
    with mapped_signal_files.MappedSignalReader(fn) as msr:
        back = msr.get_alphabet_information()
        assert (back.alphabet, back.collapse_alphabet, back.mod_long_names) == (
            'ACGTZY', 'ACGTCA', ['5mC', '6mA'])
        assert len(msr.batch_names) == 4 and msr.get_read_ids() == [r.read_id for r in reads]
        for a, b in zip(msr.reads(), reads):
            np.testing.assert_array_equal(a.Dacs, b.Dacs)
            np.testing.assert_array_equal(a.Ref_to_signal, b.Ref_to_signal)
            np.testing.assert_array_equal(a.Reference, b.Reference)
            assert a.get_read_dictionary().keys() == b.get_read_dictionary().keys()
    with pytest.raises(NotImplementedError):
        mapped_signal_files.HDF5Writer(fn, ai, batch_format=False)


def test_train_entry_point_loads_written_file(g, tmp_path):
    """The file MappedSignalWriter writes (what bin/prepare_mapped_reads.py produces) is what
    bin/train_flipflop.py's load_data reads: same reads, chunk sampling works on them."""
    import argparse
    import importlib
    import sys
    from taiyaki_b200 import chunk_selection, mapped_signal_files
    from taiyaki_b200.alphabet import AlphabetInfo
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'bin'))
    tf = importlib.import_module('train_flipflop')
    reads = golden_reads(g)
    fn = str(tmp_path / 'mapped.hdf5')
    with mapped_signal_files.MappedSignalWriter(fn, AlphabetInfo('ACGT', 'ACGT')) as msw:
        for r in reads:
            msw.write_read(r.get_read_dictionary())
    log = type('L', (), {'write': lambda self, m: None})()
    args = argparse.Namespace(input=fn, limit=5, input_strand_list=None, mod_factor=[8.0, 1.0, 50000])
    loaded, ai, _ = tf.load_data(args, log, None)
    assert [r.read_id for r in loaded] == [r.read_id for r in reads[:5]] and ai.nbase == 4
    np.random.seed(17)
    fp = chunk_selection.sample_filter_parameters(loaded, 30, 1000, 10.0, 10.0, 0.1, 5, 1.1)
    chunks, rej = chunk_selection.sample_chunks(loaded, 10, 1000, fp)
    assert len(chunks) == 10 and all(c.sig_len == 1000 for c in chunks)


def test_hdf5_errors(tmp_path):
    from taiyaki_b200 import hdf5_min
    p = tmp_path / 'x.hdf5'
    p.write_bytes(b'not hdf5 at all' * 10)
    with pytest.raises(hdf5_min.Hdf5FormatError):
        hdf5_min.File(str(p))
    p.write_bytes(b'\x89HDF\r\n\x1a\n' + bytes([2]) + bytes(200))
    with pytest.raises(hdf5_min.Hdf5FormatError, match='superblock version 2'):
        hdf5_min.File(str(p))


def test_hdf5_chunk_filters():
    """The shuffle + deflate pipeline of the batched format's datasets, undone in reverse."""
    import zlib
    from taiyaki_b200 import hdf5_min
    data = np.arange(-300, 700, dtype='<i2')
    shuffled = np.frombuffer(data.tobytes(), dtype='u1').reshape(-1, 2).T.tobytes()
    raw = zlib.compress(shuffled)
    ds = hdf5_min.Dataset.__new__(hdf5_min.Dataset)
    ds.filters = [(2, (2,)), (1, (4,))]
    ds.datatype = type('dt', (), {'size': 2})()
    np.testing.assert_array_equal(np.frombuffer(ds._unfilter(raw, 0), dtype='<i2'), data)
    # a chunk whose deflate stage was skipped (filter mask bit 1)
    np.testing.assert_array_equal(np.frombuffer(ds._unfilter(shuffled, 2), dtype='<i2'), data)
    ds.filters = [(32000, ())]
    with pytest.raises(hdf5_min.Hdf5FormatError):
        ds._unfilter(raw, 0)


def test_chunk_sampling_on_real_reads_matches_reference(g):
    """Same numpy seeds -> the reference's chunks: reads, start samples, currents, label
    sequences, dwell statistics, rejection counts, filter parameters."""
    from taiyaki_b200 import chunk_selection
    reads = golden_reads(g)
    n, chunk_len, fmd, fxd, fpass, stride, pbuf = g['fp_args']
    np.random.seed(int(g['chunk_seed'][0]))
    fp = chunk_selection.sample_filter_parameters(reads, int(n), int(chunk_len), fmd, fxd, fpass,
                                                  int(stride), pbuf)
    assert fp.median_meandwell == float(g['fp_median_meandwell'])
    assert fp.mad_meandwell == float(g['fp_mad_meandwell'])
    np.random.seed(int(g['chunk_seed'][1]))
    fp2 = fp._replace(filter_mean_dwell=float(g['chunk_filter'][0]),
                      filter_max_dwell=float(g['chunk_filter'][1]))
    chunks, rej = chunk_selection.sample_chunks(reads, len(g['chunk_read']), int(chunk_len), fp2)
    ids = [str(x) for x in g['read_ids']]
    assert [ids.index(c.read_id) for c in chunks] == list(g['chunk_read'])
    assert [c.start_sample for c in chunks] == list(g['chunk_start'])
    np.testing.assert_array_equal(np.stack([c.current for c in chunks]), g['chunk_current'])
    assert [len(c.sequence) for c in chunks] == list(g['chunk_seqlen'])
    np.testing.assert_array_equal(np.concatenate([c.sequence for c in chunks]), g['chunk_seq'])
    np.testing.assert_array_equal([c.mean_dwell for c in chunks], g['chunk_mean_dwell'])
    np.testing.assert_array_equal([c.max_dwell for c in chunks], g['chunk_max_dwell'])
    assert {k: rej[k] for k in sorted(rej)} == dict(zip((str(k) for k in g['rejection_keys']),
                                                       (int(v) for v in g['rejection_counts'])))
    assert rej['meandwell'] > 0 and rej['maxdwell'] > 0      # the filters did fire on real data


def test_real_label_loss_oracle(g):
    """The oracle restatement on the real label sequences against the reference C."""
    from oracle import oracle
    oracle.build(quiet=True)
    scores, seqs, seqlen = g['loss_scores'], g['loss_seqs'], g['loss_seqlen']
    mv, st = oracle.build_indices(seqs, seqlen, 4)
    sc, gr = oracle.c_crf_flipflop_grad(scores, mv, st, seqlen, 'f32')
    np.testing.assert_allclose(sc, g['loss_score'], rtol=1e-5)
    np.testing.assert_allclose(gr, g['loss_grad'], rtol=1e-4, atol=2e-6)


# ------------------------------------------------------------------ GPU

@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


@pytest.mark.gpu
def test_crf_kernels_on_real_labels(g, dev):
    from taiyaki_b200 import ctc
    scores, seqs, seqlen = g['loss_scores'], g['loss_seqs'], g['loss_seqlen']
    nblk = scores.shape[0]
    x = torch.tensor(scores, device=dev, requires_grad=True)
    cost = ctc.crf_flipflop_loss(x, torch.tensor(seqs), torch.tensor(seqlen), 1.0)
    cost.sum().backward()
    np.testing.assert_allclose(cost.detach().cpu().numpy(), -g['loss_score'] / nblk, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), -g['loss_grad'] / nblk, rtol=1e-4, atol=2e-6 / nblk)


@pytest.mark.gpu
@pytest.mark.parametrize('T,N,reverse', [(1000, 16, False), (4000, 8, True)])
def test_device_batching_on_real_reads(g, dev, T, N, reverse):
    """Real dwell distributions (long stays, mapping gaps) through the device batch assembly:
    same chunks, labels and rejection counts as the host path on the same candidates."""
    from .test_gpu_batching import host_batch
    from taiyaki_b200 import chunk_selection, training
    from taiyaki_b200.device_batching import DeviceReadStore
    reads = golden_reads(g)
    np.random.seed(5 + T)
    md = training.NETWORK_METADATA(reverse, True, False)
    fp = chunk_selection.sample_filter_parameters(reads, 50, T, 1.5, 6.0, 0.1, 5, 1.1)
    store = DeviceReadStore(reads, dev)
    cands = store.draw_candidates(int(N / 0.1), T)
    cur, seqs, seqlens, mods, rej = host_batch(reads, cands, N, T, fp, md, 4)
    indata, dseqs, dlens, dmods, n_acc, drej = store.sample(N, T, fp, md, 4, candidates=cands)
    torch.cuda.synchronize()
    assert n_acc == cur.shape[1] == len(seqlens) and drej == rej
    np.testing.assert_array_equal(dlens.cpu().numpy(), seqlens)
    np.testing.assert_array_equal(dseqs.cpu().numpy(), seqs)
    np.testing.assert_allclose(indata[:, :, 0].cpu().numpy(), cur, rtol=2e-6, atol=2e-6)


@pytest.mark.gpu
def test_training_on_real_reads(g, dev):
    """A few optimiser steps on chunks of the real reads: finite, decreasing loss."""
    from taiyaki_b200 import chunk_selection, device_batching, helpers, training
    from taiyaki_b200.alphabet import AlphabetInfo
    root = os.path.dirname(HERE)
    np.random.seed(0)
    torch.manual_seed(0)
    ai = AlphabetInfo('ACGT', 'ACGT')
    net = helpers.load_model(os.path.join(root, 'models', 'mLstm_flipflop.py'),
                             model_metadata={'reverse': False, 'standardize': True},
                             size=64, stride=5, winlen=19, insize=1, alphabet_info=ai).to(dev)
    net_info = training.NETWORK_INFO(net=net, net_clone=None,
                                     metadata=training.parse_network_metadata(net), stride=5)
    reads = golden_reads(g)
    fp = chunk_selection.sample_filter_parameters(reads, 50, 1000, 10.0, 10.0, 0.1, 5, 1.1)
    store = device_batching.DeviceReadStore(reads, dev)
    step = training.TrainStep(net_info, torch.optim.AdamW(net.parameters(), lr=2e-3, eps=1e-6))
    losses = []
    for _ in range(30):
        gen = device_batching.prepare_random_batches(store, 1000, 16, 1, ai, fp, net_info, None)
        _, loss, _ = step(gen, sharpen=1.0)
        assert np.isfinite(loss)
        losses.append(loss)
    assert np.mean(losses[-5:]) < np.mean(losses[:5])
