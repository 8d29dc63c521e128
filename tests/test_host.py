"""CPU tests of the host-side surface: batching (chunk_selection /
signal_mapping / prepare_random_batches), flip-flop coding, the flat-gradient
all-reduce over gloo with world_size 2, the CLI parser, and bench.py's
reference arm plumbing."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flipflopfings_mirror_matches_golden(kat):
    from taiyaki_b200 import flipflopfings as fff
    np.testing.assert_array_equal(fff.flipflop_code(kat['code_in0']), kat['code_out0'])
    c = fff.flipflop_code(kat['code_in1'])
    np.testing.assert_array_equal(c, kat['code_out1'])
    np.testing.assert_array_equal(fff.move_indices(c), kat['code_move1'])
    np.testing.assert_array_equal(fff.stay_indices(c), kat['code_stay1'])
    assert fff.nstate_flipflop(4) == 40 and fff.nbase_flipflop(40) == 4
    with pytest.raises(AssertionError):
        fff.nbase_flipflop(45)


def test_synthetic_reads_and_chunk_sampling():
    from taiyaki_b200 import chunk_selection, signal_mapping
    np.random.seed(0)
    reads = signal_mapping.synthetic_reads(6, seed=3)
    r = reads[0]
    assert r.Ref_to_signal[0] == 0 and r.Ref_to_signal[-1] == len(r.Dacs)
    assert np.all(np.diff(r.Ref_to_signal) >= 1)
    fp = chunk_selection.sample_filter_parameters(reads, 40, 2000, 3.0, 10.0, 0.5, 5, 1.1)
    assert 8.0 < fp.median_meandwell < 10.0
    chunks, rej = chunk_selection.sample_chunks(reads, 12, 2000, fp)
    assert len(chunks) == 12 and rej['pass'] == 12
    for c in chunks:
        assert c.sig_len == 2000 and 150 < c.seq_len < 300
        assert abs(float(np.mean(c.current))) < 0.5
    # too-short reads are rejected with the reference's reason string
    short = reads[0].get_chunk_with_sample_length(10 ** 7)
    assert short.reject_reason == 'tooshort' and not short.accepted
    # the path-buffer filter (signal_mapping.py:699-703)
    tight = fp._replace(path_buffer=100.0)
    ch = reads[0].get_chunk_with_sample_length(2000)
    ch.apply_filters(tight)
    assert ch.reject_reason == 'pathbuffer'


def test_prepare_random_batches_layout():
    from taiyaki_b200 import chunk_selection, flipflopfings, signal_mapping, training
    from taiyaki_b200.alphabet import AlphabetInfo
    np.random.seed(1)
    reads = signal_mapping.synthetic_reads(5, seed=4)
    fp = chunk_selection.sample_filter_parameters(reads, 30, 1000, 3.0, 10.0, 0.5, 5, 1.1)
    md = training.NETWORK_METADATA(False, True, False)
    net_info = training.NETWORK_INFO(net=None, net_clone=None, metadata=md, stride=5)
    gen = training.prepare_random_batches(reads, 1000, 7, 2, AlphabetInfo('ACGT', 'ACGT'), fp,
                                          net_info, None, pin=False)
    batches = list(gen)
    assert len(batches) == 2
    indata, seqs, seqlens, mod_cats, n, rej = batches[0]
    assert indata.shape == (1000, 7, 1) and indata.dtype == torch.float32
    assert seqs.dtype == torch.long and int(seqlens.sum()) == len(seqs) and mod_cats is None
    s0 = seqs[:int(seqlens[0])].numpy()
    assert s0.max() < 8
    # flip-flop coded: no two equal consecutive codes
    assert np.all(s0[1:] != s0[:-1])
    assert flipflopfings.stay_indices(s0).max() < 40


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from taiyaki_b200.training import FlatGradients
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    net[2].bias.requires_grad = False            # a frozen parameter, like bias_hh
    flat = FlatGradients(net.parameters())
    torch.manual_seed(100 + rank)                # different data per rank
    x = torch.randn(6, 5)
    flat.zero()
    net(x).pow(2).mean().backward()
    assert flat.check_views()
    local = flat.flat.clone()
    flat.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered) / world
    ok = torch.allclose(flat.flat, want, atol=1e-7)
    maxs = flat.grad_maxs()
    ok = ok and maxs.numel() == 3 and net[2].bias.grad is None
    ok = ok and torch.allclose(maxs[0], net[0].weight.grad.abs().max())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    mgr = ctx.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0] and out[1]


def test_rolling_mad_and_med_mad():
    from taiyaki_b200 import maths
    rm = maths.RollingMAD(2, n_mads=1.0, window=5)
    for i in range(4):
        assert rm.update([1.0 + i, 10.0]) is None
    thr = rm.update([5.0, 10.0])
    assert thr.shape == (2,) and abs(thr[1] - 10.0) < 1e-6 and thr[0] > 3.0
    med, mad = maths.med_mad(np.array([1.0, 2.0, 3.0, 4.0, 100.0]))
    assert med == 3.0 and abs(mad - 1.4826) < 1e-6


def test_train_cli_parser_defaults_match_reference():
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    import importlib
    tf = importlib.import_module('train_flipflop')
    a = tf.get_train_flipflop_parser().parse_args(['models/mLstm_flipflop.py', 'synthetic:10'])
    # bin/_bin_argparse.py:9-208
    assert (a.size, a.stride, a.winlen) == (384, 5, 19)
    assert (a.chunk_len_min, a.chunk_len_max, a.min_sub_batch_size) == (3000, 8000, 128)
    assert a.lr_max == 4e-3 and a.lr_min == 1e-4 and a.niteration == 150000
    assert tuple(a.sharpen) == (1.0, 1.0, 25000) and a.gradient_clip_num_mads == 0
    assert a.weight_decay == 0.01 and a.eps == 1e-6 and a.warmup_batches == 200
    assert a.mod_prior_factor is None and a.num_mod_weight_reads == 5000


def test_boolean_flag_pairs_parse_like_the_reference():
    """taiyaki/cmdargs.py:90-127 AutoBool: `--flag` / `--no-flag`, neither takes a value -- so the
    reference's command lines, where such a flag may stand right before the positional arguments
    (`basecall.py --fastq reads/ model`), parse the same here."""
    import argparse
    import importlib
    from taiyaki_b200.cmdargs import AutoBool
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    sys.path.insert(0, os.path.join(ROOT, 'misc'))
    p = argparse.ArgumentParser()
    p.add_argument('--thing', default=True, action=AutoBool, help='A thing')
    p.add_argument('where')
    assert p.parse_args(['x']).thing is True and p.parse_args(['--no-thing', 'x']).thing is False
    assert p.parse_args(['--thing', 'x']).where == 'x' and '(Default: --thing)' in p.format_help()
    with pytest.raises(ValueError):
        argparse.ArgumentParser().add_argument('--nodefault', action=AutoBool)
    bc = importlib.import_module('basecall').get_parser()
    a = bc.parse_args(['--fastq', 'reads/', 'model.checkpoint'])
    assert a.fastq is True and (a.input_folder, a.model) == ('reads/', 'model.checkpoint')
    a = bc.parse_args(['--no-posterior', '--reverse', '--no-recursive', '--quiet', 'reads/', 'm'])
    assert (a.posterior, a.reverse, a.recursive, a.quiet, a.fastq) == (False, True, False, True, False)
    tf = importlib.import_module('train_flipflop').get_train_flipflop_parser()
    a = tf.parse_args(['--no-standardize', '--reverse', '--full_filter_status', '--overwrite', 'model.py', 'in.hdf5'])
    assert (a.standardize, a.reverse, a.full_filter_status, a.overwrite, a.quiet) == (False, True, True, True, False)
    a = tf.parse_args(['--quiet', 'model.py', 'in.hdf5'])
    assert a.quiet is True and a.standardize is True and a.model == 'model.py'
    pm = importlib.import_module('prepare_mapped_reads').get_parser()
    a = pm.parse_args(['--overwrite', 'reads/', 'p.tsv', 'out.hdf5', 'model', 'refs.fa'])
    assert a.overwrite is True and a.recursive is True and a.input_folder == 'reads/'
    gp = importlib.import_module('generate_per_read_params').get_parser()
    assert gp.parse_args(['--no-recursive', 'reads/']).recursive is False
    gr = importlib.import_module('get_refs_from_sam').get_parser()
    a = gr.parse_args(['--reverse', '--complement', 'genome.fa', 'a.sam', 'b.sam'])
    assert (a.reverse, a.complement, a.reference, a.input) == (True, True, 'genome.fa', ['a.sam', 'b.sam'])
    mg = importlib.import_module('merge_mappedsignalfiles').get_parser()
    assert mg.parse_args(['out', '--input', 'a', 'None', '--no-load_in_mem']).load_in_mem is False


@pytest.mark.skipif(not os.path.isdir('/root/reference/test/unit'),
                    reason='the reference tree is only in the build container')
@pytest.mark.parametrize('name,ntests', [('test_cmdargs', 13), ('test_iterate_fast5_reads', 9), ('test_maths', 4)])
def test_reference_unit_tests_pass_against_this_package(name, ntests):
    """The reference's OWN unit tests of the host-side modules (test/unit/<name>.py, loaded from
    where they lie, unmodified) with `taiyaki` resolving to this repository's alias package: what a
    maintainer switching the import path would run first.  (The tests of the device operators --
    test_ctc_loss, test_decode, test_flipflop_remap, test_layers -- need a GPU, and the GPU box has
    no reference tree; tests/test_gpu_*.py restate their cases.)"""
    import importlib.util
    import io
    import types
    import unittest
    import taiyaki
    assert os.path.dirname(os.path.dirname(os.path.abspath(taiyaki.__file__))) == ROOT
    pkg = types.ModuleType('reference_unit_tests')       # stands in for test/unit/__init__.py
    pkg.__path__ = []
    pkg.DATA_DIR = '/root/reference/test/data'
    sys.modules['reference_unit_tests'] = pkg
    spec = importlib.util.spec_from_file_location('reference_unit_tests.' + name,
                                                  '/root/reference/test/unit/%s.py' % name)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    state = np.random.get_state()                        # test_maths seeds the global generator
    try:
        result = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(
            unittest.defaultTestLoader.loadTestsFromModule(module))
    finally:
        np.random.set_state(state)
    assert result.testsRun == ntests and result.wasSuccessful(), result.failures + result.errors


@pytest.mark.skipif(not os.path.isdir('/root/reference/models'),
                    reason='the reference tree is only in the build container')
def test_dump_json_of_the_shipped_checkpoints_matches_the_reference(tmp_path):
    """bin/dump_json.py on the reference's shipped checkpoints: the text equals, byte for byte, what the
    reference's own layer classes and JsonEncoder produce from the same files (md5 and length in
    tests/golden/model_json.npz, make_golden.py model_json) -- layer descriptions, Guppy gate order of
    the GRU parameters, every weight."""
    import hashlib
    import importlib
    from taiyaki_b200 import helpers
    from taiyaki_b200.json import JsonEncoder
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    cli = importlib.import_module('dump_json')
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'model_json.npz'))
    name = 'mGru_flipflop_remapping_model_r9_DNA'
    out = tmp_path / 'model.json'
    cli.main(['--output', str(out), '/root/reference/models/%s.checkpoint' % name])
    text = out.read_text()
    assert [hashlib.md5(text.encode()).hexdigest(), str(len(text))] == list(gold[name][:2])
    parsed = json.loads(text)
    assert parsed['md5sum'] == gold[name][2] and parsed['type'] == 'serial' and len(parsed['sublayers']) == 7
    with pytest.raises(RuntimeError):                   # FileAbsent: an existing output is not replaced
        cli.main(['--output', str(out), '/root/reference/models/%s.checkpoint' % name])
    name = 'mLstm_flipflop_model_r941_DNA'              # 58 MB of text: compared without indentation
    fn = '/root/reference/models/%s.checkpoint' % name
    json_out = helpers.load_model(fn).json()
    json_out['md5sum'] = cli.file_md5(fn)
    text = json.dumps(json_out, cls=JsonEncoder)
    assert [hashlib.md5(text.encode()).hexdigest(), str(len(text)), json_out['md5sum']] == list(gold[name])


def test_dump_json_cli_on_a_saved_model(tmp_path, capsys):
    """dump_json.py on a checkpoint written by helpers.save_model: valid JSON on stdout, the layer
    descriptions of the model and its parameters in Guppy's layout."""
    import importlib
    from taiyaki_b200 import helpers
    from taiyaki_b200.alphabet import AlphabetInfo
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    cli = importlib.import_module('dump_json')
    state = np.random.get_state()
    try:
        np.random.seed(5)
        net = helpers.load_model(os.path.join(ROOT, 'models', 'mGru_flipflop.py'), size=32, stride=2,
                                 winlen=19, insize=1, alphabet_info=AlphabetInfo('ACGT', 'ACGT'))
    finally:
        np.random.set_state(state)
    ckpt, _ = helpers.save_model(net, str(tmp_path))
    capsys.readouterr()
    cli.main([ckpt])
    parsed = json.loads(capsys.readouterr().out)
    assert parsed['md5sum'] == cli.file_md5(ckpt)
    kinds = [layer['type'] for layer in parsed['sublayers']]
    assert kinds == ['convolution', 'reverse', 'GruMod', 'reverse', 'GruMod', 'reverse', 'GlobalNormTwoState']
    gru = parsed['sublayers'][2]
    assert (gru['size'], gru['insize'], gru['bias']) == (32, 32, True)
    w = net.sublayers[2].cudnn_gru.weight_ih_l0.detach().numpy().reshape(3, 32, 32)     # cuDNN order r, z, n
    np.testing.assert_array_equal(np.array(gru['params']['iW'], dtype='f4'), w[[1, 0, 2]])   # Guppy: z, r, n
    conv = parsed['sublayers'][0]
    assert (conv['winlen'], conv['stride'], conv['padding'], conv['activation']) == (19, 2, [9, 9], 'tanh')
    np.testing.assert_array_equal(np.array(conv['params']['W'], dtype='f4'),
                                  net.sublayers[0].conv.weight.detach().numpy())


def test_argument_types_follow_the_reference():
    """taiyaki/cmdargs.py types, on the value lists of the reference's test/unit/test_cmdargs.py."""
    import argparse
    from taiyaki_b200 import cmdargs
    eps = sys.float_info.epsilon

    def refuses(f, values):
        for x in values:
            with pytest.raises(argparse.ArgumentTypeError):
                f(x)
    for x in [1e-30, eps, 1e-5, 1.0, 1e5, 1e30]:
        assert cmdargs.Positive(float)(x) == x
    refuses(cmdargs.Positive(float), [-1.0, -eps, -1e-5, 0.0])
    assert [cmdargs.Positive(int)(x) for x in [1, 10, 10000]] == [1, 10, 10000]
    refuses(cmdargs.Positive(int), [-1, 0])
    for x in [1e-30, eps, 1e-5, 0.0, 1.0, 1e5, 1e30]:
        assert cmdargs.NonNegative(float)(x) == x
    refuses(cmdargs.NonNegative(float), [-1.0, -eps, -1e-5])
    assert [cmdargs.NonNegative(int)(x) for x in [0, 1, 10, 10000]] == [0, 1, 10, 10000]
    refuses(cmdargs.NonNegative(int), [-1, -10])
    for x in [1e-30, eps, 1e-5, 0.0, 1.0, 1.0 - 1e-5, 1.0 - eps, 1.0 - 1e-30]:
        assert cmdargs.proportion(x) == x
    refuses(cmdargs.proportion, [-1e-30, -eps, -1e-5, 1.0 + 1e-5, 1.0 + eps])
    assert [cmdargs.Bounded(int, 0, 10)(x) for x in range(11)] == list(range(11))
    refuses(cmdargs.Bounded(int, 0, 10), [-2, -1, 11, 12])
    assert cmdargs.Maybe(cmdargs.Positive(int))('None') is None and cmdargs.Maybe(int)('7') == 7
    refuses(cmdargs.Maybe(cmdargs.Positive(int)), ['0', 'seven'])
    p = argparse.ArgumentParser()
    p.add_argument('device', action=cmdargs.DeviceAction)
    p.add_argument('--table', action=cmdargs.FileExists)
    p.add_argument('--fresh', action=cmdargs.FileAbsent)
    assert [p.parse_args([d]).device for d in ('2', 'cuda2', 'cuda:2', 'cuda', 'cpu')] == [2, 2, 'cuda:2', 'cuda', 'cpu']
    assert p.parse_args(['0', '--table', __file__]).table == __file__
    with pytest.raises(RuntimeError):
        p.parse_args(['0', '--table', __file__ + '.absent'])
    with pytest.raises(RuntimeError):
        p.parse_args(['0', '--fresh', __file__])
    # the entry points use them: a value the reference refuses is refused here
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    import importlib
    tf = importlib.import_module('train_flipflop').get_train_flipflop_parser()
    for bad in (['--size', '0'], ['--lr_max', '-1'], ['--filter_path_buffer', '0.9'], ['--seed', '0'],
                ['--chunk_len_min', '-5'], ['--sub_batches', '0']):
        with pytest.raises(SystemExit):
            tf.parse_args(bad + ['model.py', 'in.hdf5'])
    a = tf.parse_args(['--gradient_clip_num_mads', 'None', '--filter_max_dwell', 'None', '--device', '3',
                       'model.py', 'in.hdf5'])
    assert a.gradient_clip_num_mads is None and a.filter_max_dwell is None and a.device == 3


def test_train_cli_mod_prior_factor():
    """--mod_prior_factor: prior odds of the sampled reads raised to the factor
    (train_flipflop.py:312-326); without the flag every category weighs 1."""
    import io
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    import importlib
    tf = importlib.import_module('train_flipflop')
    from taiyaki_b200 import signal_mapping
    from taiyaki_b200.alphabet import AlphabetInfo
    info = AlphabetInfo('ACGTZ', 'ACGTC', ['5mC'])
    reads = signal_mapping.synthetic_reads(20, seed=7, mod_fraction=0.5)
    parser = tf.get_train_flipflop_parser()
    plain = parser.parse_args(['models/mGru_cat_mod_flipflop.py', 'synthetic:20:5mC'])
    np.testing.assert_array_equal(tf.mod_prior_weights(plain, info, reads, io.StringIO()), np.ones(5, 'f4'))
    log = io.StringIO()
    args = parser.parse_args(['--mod_prior_factor', '0.5', 'models/mGru_cat_mod_flipflop.py', 'synthetic:20:5mC'])
    w = tf.mod_prior_weights(args, info, reads, log)
    np.random.seed(0)
    expect = np.power(info.compute_log_odds_weights(reads, 5000), 0.5)
    np.testing.assert_allclose(w, expect, rtol=1e-6)
    assert w.dtype == np.float32 and w[1] > 0 and w[2] > 0
    assert 'Computed modbase log odds priors' in log.getvalue() and 'Applied mod_prior_factor' in log.getvalue()


def test_model_definition_files_build():
    from taiyaki_b200 import helpers, layers
    from taiyaki_b200.alphabet import AlphabetInfo
    np.random.seed(0)
    net = helpers.load_model(os.path.join(ROOT, 'models', 'mLstm_flipflop.py'), size=64,
                             stride=5, winlen=19, insize=1,
                             alphabet_info=AlphabetInfo('ACGT', 'ACGT'))
    assert isinstance(net, layers.Serial) and len(net.sublayers) == 9
    assert isinstance(net.sublayers[3], layers.Reverse)
    names = [n for n, _ in net.named_parameters()]
    assert 'sublayers.3.layer.lstm.weight_hh_l0' in names
    cm = helpers.load_model(os.path.join(ROOT, 'models', 'mGru_cat_mod_flipflop.py'), size=64,
                            stride=2, winlen=19, insize=1,
                            alphabet_info=AlphabetInfo('ACGTZ', 'ACGTC', ['5mC']))
    last = cm.sublayers[-1]
    assert layers.is_cat_mod_model(cm) and last.size == 42   # linear outputs; forward emits 45
    np.testing.assert_array_equal(last.can_mods_offsets, [0, 1, 3, 4, 5])
    # full-size parameter count of the reference's mLstm_flipflop
    big = helpers.load_model(os.path.join(ROOT, 'models', 'mLstm_flipflop.py'), size=256,
                             stride=5, winlen=19, insize=1,
                             alphabet_info=AlphabetInfo('ACGT', 'ACGT'))
    assert sum(p.numel() for p in big.parameters()) == 2720400
    assert sum(p.numel() for p in big.parameters() if p.requires_grad) == 2715280


def test_mod_prior_weights():
    """Modified-base prior weights (--mod_prior_factor) against the reference's
    AlphabetInfo.compute_log_odds_weights / compute_mod_inv_freq_weights (alphabet.py:35-100),
    golden values from tests/golden/make_golden.py mod_weights."""
    from taiyaki_b200.alphabet import AlphabetInfo
    from taiyaki_b200.signal_mapping import SignalMapping
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'mod_weights.npz'))
    for alpha, collapse, names in [('ACGTZ', 'ACGTC', ['5mC']), ('ACGTZY', 'ACGTCA', ['5mC', '6mA']),
                                   ('ACGTZYX', 'ACGTCAC', ['5mC', '6mA', '5hmC'])]:
        rng = np.random.RandomState(11)
        refs = [rng.randint(0, len(alpha), size=200 + 13 * i).astype(np.int16) for i in range(7)]
        info = AlphabetInfo(alpha, collapse, names)
        reads = [SignalMapping(np.zeros(4, 'i2'), np.arange(len(r) + 1), r) for r in refs]
        np.testing.assert_array_equal(info.compute_log_odds_weights(reads, 100), gold[alpha + '_log_odds'])
        np.testing.assert_array_equal(info.compute_mod_inv_freq_weights([{'Reference': r} for r in refs], 100),
                                      gold[alpha + '_inv_freq'])
        # a sample smaller than the read set is a draw without replacement from it
        np.random.seed(3)
        w = info.compute_log_odds_weights(reads, 3)
        assert w.shape == (len(alpha),) and np.isfinite(w).all()
    # a label that never occurs has no ratio: same refusal as the reference
    info = AlphabetInfo('ACGTZ', 'ACGTC', ['5mC'])
    with pytest.raises(NotImplementedError):
        info.compute_log_odds_weights([{'Reference': np.array([0, 1, 2, 3], 'i2')}], 10)


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '1', '--ref-chunks', '1'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port')
    assert line['e2e']['h2d_bytes_per_step'] == 0
