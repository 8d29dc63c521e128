#!/usr/bin/env python3
"""Train a flip-flop neural network -- entry point kept from taiyaki's
bin/train_flipflop.py (same positional arguments and flags, same logs:
model.log / batch.log / validation.log, same checkpoint names), running the
hot path on the B200-native kernels of taiyaki_b200.

    train_flipflop.py [flags] model.py input

`input` is a mapped-signal source: a mapped-signal HDF5 file, per-read or batched
(taiyaki/mapped_signal_files.py; read by taiyaki_b200/mapped_signal_files.py over a
plain-Python HDF5 decoder, this image has no h5py), or `synthetic:NREADS[:5mC]`,
which generates r9.4.1-like reads in memory (taiyaki_b200/signal_mapping.py).  Both
go through the same chunk_selection / prepare_random_batches path.

Multi-GPU: one process per GPU, `torchrun --nproc-per-node G bin/train_flipflop.py ...`
(LOCAL_RANK from the environment, or --local_rank as in the reference,
_bin_argparse.py:155-156).  Gradients are averaged with one NCCL all-reduce of
a flat buffer per optimiser step (taiyaki_b200/training.py: FlatGradients).
"""
import argparse
import math
import os
import sys
import time
from collections import defaultdict, namedtuple
from itertools import islice
from shutil import copyfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import (chunk_selection, device_batching, helpers, layers,  # noqa: E402
                          mapped_signal_files, maths, signal_mapping, training)
from taiyaki_b200.alphabet import AlphabetInfo  # noqa: E402
from taiyaki_b200.cmdargs import (AutoBool, Bounded, DeviceAction, FileExists, Maybe,  # noqa: E402
                                  NonNegative, Positive)

DOTROWLENGTH = 50
MODEL_LOG_FILENAME, BATCH_LOG_FILENAME, VAL_LOG_FILENAME = 'model.log', 'batch.log', 'validation.log'
BATCH_HEADER = '\t'.join(('iter', 'loss', 'gradientmax', 'gradientcap', 'learning_rate',
                          'chunk_len')) + '\n'
BATCH_TMPLT = '{}\t{:5.3f}\t{}\t{}\t{:.2e}\t{}\n'
VAL_HEADER = 'iter\tloss\n'
VAL_TMPLT = '{}\t{:5.3f}\n'
MAIN_LOG_POLKA_TMPLT = (' {:5d} {:7.5f}   {:5.2f}s ({:.2f} ksample/s {:.2f} kbase/s) lr={:.2e}')
MAIN_LOG_VAL_TMPLT = ('iteration: {} validation_loss: {:7.5f} ({:5.2} Mbase in {:5.2f} s, '
                      '{:.2f} kbase/s)\n')

RESOURCE_INFO = namedtuple('RESOURCE_INFO', ('is_multi_gpu', 'is_lead_process', 'device'))
OPTIM_INFO = namedtuple('OPTIM_INFO', ('optimiser', 'lr_warmup', 'lr_scheduler', 'rolling_mads'))
TRAIN_PARAMS = namedtuple('TRAIN_PARAMS', (
    'niteration', 'sharpen', 'chunk_len_min', 'chunk_len_max', 'min_sub_batch_size',
    'sub_batches', 'save_every', 'outdir', 'full_filter_status', 'host_batching'))
SHARPEN = namedtuple('SHARPEN', ('min', 'max', 'niter'))
MOD_FACTOR = namedtuple('MOD_FACTOR', ('start', 'final', 'niter'))
LOGS = namedtuple('LOGS', ('main', 'batch', 'validation'))
LOGS.__new__.__defaults__ = (None, None)


def get_train_flipflop_parser():
    """Flags of bin/_bin_argparse.py:9-208 that reach the training path."""
    p = argparse.ArgumentParser(description='Train flip-flop neural network',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    g = p.add_argument_group('Model Arguments')
    g.add_argument('--size', default=384, type=Positive(int), metavar='neurons')
    g.add_argument('--stride', default=5, type=Positive(int), metavar='samples')
    g.add_argument('--winlen', default=19, type=Positive(int))
    g = p.add_argument_group('Training Arguments')
    g.add_argument('--adam', nargs=2, default=[0.9, 0.999], type=NonNegative(float), metavar=('beta1', 'beta2'))
    g.add_argument('--eps', default=1e-6, type=Positive(float))
    g.add_argument('--niteration', default=150000, type=Positive(int))
    g.add_argument('--weight_decay', default=0.01, type=NonNegative(float))
    g.add_argument('--gradient_clip_num_mads', default=0, type=Maybe(NonNegative(float)))
    g.add_argument('--lr_max', default=4.0e-3, type=Positive(float))
    g.add_argument('--lr_min', default=1.0e-4, type=Positive(float))
    g.add_argument('--lr_warmup', default=None, type=Positive(float))
    g.add_argument('--min_momentum', default=None, type=Positive(float))
    g.add_argument('--seed', default=None, type=Positive(int))
    g.add_argument('--sharpen', default=(1.0, 1.0, 25000), nargs=3, type=float,
                   metavar=('min', 'max', 'niter'))
    g.add_argument('--warmup_batches', type=int, default=200)
    g = p.add_argument_group('Data Arguments')
    g.add_argument('--filter_max_dwell', default=10.0, type=Maybe(Positive(float)))
    g.add_argument('--filter_mean_dwell', default=3.0, type=Maybe(Positive(float)))
    g.add_argument('--filter_min_pass_fraction', default=0.5, type=Maybe(Positive(float)))
    g.add_argument('--filter_path_buffer', default=1.1, type=Bounded(float, lower=1.0))
    g.add_argument('--limit', default=None, type=Maybe(Positive(int)))
    g.add_argument('--input_strand_list', default=None, action=FileExists)
    g.add_argument('--reverse', default=False, action=AutoBool, help='Reverse input sequence and current')
    g.add_argument('--sample_nreads_before_filtering', type=NonNegative(int), default=100000)
    g.add_argument('--chunk_len_min', default=3000, type=Positive(int))
    g.add_argument('--chunk_len_max', default=8000, type=Positive(int))
    g.add_argument('--min_sub_batch_size', default=128, type=Positive(int))
    g.add_argument('--reporting_sub_batches', default=100, type=Positive(int))
    g.add_argument('--standardize', default=True, action=AutoBool,
                   help='Standardize currents for each read')
    g.add_argument('--sub_batches', default=1, type=Positive(int))
    g = p.add_argument_group('Compute Arguments')
    g.add_argument('--device', default='cuda:0', action=DeviceAction,
                   help='GPU to use: an integer, "cuda:2", "cuda2" or "cuda" (this path has no CPU mode)')
    g.add_argument('--local_rank', type=int, default=None, help=argparse.SUPPRESS)
    g = p.add_argument_group('Output Arguments')
    g.add_argument('--full_filter_status', default=False, action=AutoBool,
                   help='Output full chunk filtering statistics')
    g.add_argument('--outdir', default='training')
    g.add_argument('--overwrite', default=False, action=AutoBool, help='Whether to overwrite any output files')
    g.add_argument('--quiet', default=False, action=AutoBool, help="Don't print progress information to stdout")
    g.add_argument('--save_every', type=Positive(int), default=2500)
    g.add_argument('--cuda_graphs', default=False, action='store_true',
                   help='Replay forward + loss + backward as a CUDA graph once a batch shape has been '
                        'seen three times (useful with --chunk_len_min == --chunk_len_max and short '
                        'chunks, where the step is bound by kernel launches; same results)')
    g.add_argument('--host_batching', default=False, action='store_true',
                   help='Assemble training batches with the numpy path of the reference '
                        'instead of on the device (taiyaki_b200.device_batching)')
    g = p.add_argument_group('Modified Base Arguments')
    g.add_argument('--mod_factor', default=(8.0, 1.0, 50000), nargs=3, type=float,
                   metavar=('start', 'final', 'niter'))
    g.add_argument('--mod_prior_factor', type=float, default=None,
                   help='Exponent applied to the prior weights of the modified-base categories '
                        'estimated from the training reads.  Default: no prior (all weights 1)')
    g.add_argument('--num_mod_weight_reads', type=int, default=5000,
                   help='Number of reads sampled to estimate the modified-base prior weights')
    p.add_argument('model', help='File to read python model (or checkpoint) from')
    p.add_argument('input', help='Mapped-signal HDF5 file, or synthetic:NREADS[:5mC]')
    return p


def parse_init_args(args):
    """train_flipflop.py:215-280"""
    local_rank = args.local_rank
    if local_rank is None and 'LOCAL_RANK' in os.environ and \
            int(os.environ.get('WORLD_SIZE', '1')) > 1:
        local_rank = int(os.environ['LOCAL_RANK'])     # torchrun
    args.local_rank = local_rank
    is_multi_gpu = local_rank is not None
    is_lead_process = (not is_multi_gpu) or local_rank == 0
    seed = (np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32)
            if args.seed is None else args.seed)
    main_log_fn = os.path.join(args.outdir, MODEL_LOG_FILENAME)
    if is_lead_process:
        if os.path.exists(args.outdir) and not args.overwrite:
            sys.stderr.write('Error: {} exists but --overwrite is false\n'.format(args.outdir))
            sys.exit(1)
        os.makedirs(args.outdir, exist_ok=True)
        if args.model.endswith('.py'):
            copyfile(args.model, os.path.join(args.outdir, 'model.py'))
        logs = LOGS(main=helpers.Logger(main_log_fn, args.quiet),
                    batch=open(os.path.join(args.outdir, BATCH_LOG_FILENAME), 'w', buffering=1),
                    validation=open(os.path.join(args.outdir, VAL_LOG_FILENAME), 'w',
                                    buffering=1))
        logs.batch.write(BATCH_HEADER)
        logs.validation.write(VAL_HEADER)
        if args.save_every % DOTROWLENGTH != 0:
            se2 = int(math.ceil(args.save_every / DOTROWLENGTH)) * DOTROWLENGTH
            logs.main.write('* --save_every {} not a multiple of {}, rounding to {}\n'.format(
                args.save_every, DOTROWLENGTH, se2))
            args.save_every = se2
        if args.chunk_len_min > args.chunk_len_max:
            raise ValueError('--chunk_len_min greater than --chunk_len_max')
        logs.main.write('* Using random seed: {}\n'.format(seed))
    if is_multi_gpu:
        try:
            torch.distributed.init_process_group(backend='nccl')
        except Exception:
            raise Exception(
                'Unable to start multiprocessing group. The most likely reason is that the '
                'script is running with local_rank set but without the set-up for distributed '
                'operation (use torchrun).')
        if not is_lead_process:
            torch.distributed.barrier()      # wait for rank 0 to create outdir
            logs = LOGS(main=helpers.Logger(main_log_fn + '.rank%d' % local_rank, True))
        else:
            torch.distributed.barrier()
        device = torch.device('cuda', local_rank)
        seed += local_rank                   # different data streams per GPU
    else:
        device = torch.device(args.device)
    if device.type != 'cuda':
        raise RuntimeError('taiyaki_b200 trains on a CUDA device only (got --device {})'.format(
            args.device))
    torch.cuda.set_device(device)
    np.random.seed(int(seed) % (2 ** 32))
    torch.manual_seed(int(seed))
    return RESOURCE_INFO(is_multi_gpu, is_lead_process, device), logs


def get_read_ids(filename):
    """read_id column of a tab-separated strand list (helpers.py:200-209)."""
    with open(filename) as fh:
        header = fh.readline().rstrip('\n').split('\t')
        col = header.index('read_id')
        return [line.rstrip('\n').split('\t')[col] for line in fh if line.strip()]


def load_data(args, log, res_info):
    """train_flipflop.py:283-329; mapped-signal HDF5 or the in-memory synthetic source."""
    log.write('* Loading data from {}\n'.format(args.input))
    if not args.input.startswith('synthetic:'):
        # mapped-signal HDF5, per-read or batched (train_flipflop.py:285-303)
        if args.input_strand_list is not None:
            read_ids = list(set(get_read_ids(args.input_strand_list)))
            log.write(('* Will train from a subset of {} strands, determined ' +
                       'by read_ids in input strand list\n').format(len(read_ids)))
        else:
            log.write('* Reads not filtered by id\n')
            read_ids = None
        if args.limit is not None:
            log.write('* Limiting number of strands to {}\n'.format(args.limit))
        with mapped_signal_files.MappedSignalReader(args.input) as msr:
            alphabet_info = msr.get_alphabet_information()
            read_data = list(islice(msr.reads(read_ids), args.limit))
        log.write('* Using alphabet definition: {}\n'.format(str(alphabet_info)))
        if len(read_data) == 0:
            log.write('* No reads remaining for training, exiting.\n')
            sys.exit(1)
        log.write('* Loaded {} reads.\n'.format(len(read_data)))
        mod_cat_weights = mod_prior_weights(args, alphabet_info, read_data, log)
        return read_data, alphabet_info, training.MOD_INFO(mod_cat_weights, MOD_FACTOR(*args.mod_factor))
    parts = args.input.split(':')
    nreads = int(parts[1])
    with_mods = len(parts) > 2 and parts[2].lower() == '5mc'
    if args.limit is not None:
        nreads = min(nreads, args.limit)
    read_data = signal_mapping.synthetic_reads(nreads, seed=7,
                                               mod_fraction=0.5 if with_mods else 0.0)
    alphabet_info = (AlphabetInfo('ACGTZ', 'ACGTC', ['5mC']) if with_mods else
                     AlphabetInfo('ACGT', 'ACGT'))
    log.write('* Using alphabet definition: {}\n'.format(str(alphabet_info)))
    log.write('* Loaded {} reads.\n'.format(len(read_data)))
    mod_cat_weights = mod_prior_weights(args, alphabet_info, read_data, log)
    mod_info = training.MOD_INFO(mod_cat_weights, MOD_FACTOR(*args.mod_factor))
    return read_data, alphabet_info, mod_info


def mod_prior_weights(args, alphabet_info, read_data, log):
    """Per-category weights of the cat-mod loss: ones, or with --mod_prior_factor the prior odds
    estimated from the reads raised to that power (train_flipflop.py:312-326)."""
    factor = getattr(args, 'mod_prior_factor', None)
    if factor is None:
        return np.ones(alphabet_info.nbase, dtype=np.float32)

    def listing(w):
        return '  '.join('{}:{:.4f}'.format(b, x) for b, x in zip(alphabet_info.alphabet, w))
    weights = alphabet_info.compute_log_odds_weights(read_data, getattr(args, 'num_mod_weight_reads', 5000))
    log.write('* Computed modbase log odds priors:  {}\n'.format(listing(weights)))
    if factor != 1.0:
        weights = np.power(weights, factor)
        log.write('* Applied mod_prior_factor to modbase log odds priors:  {}\n'.format(listing(weights)))
    return weights.astype(np.float32)


def load_network(args, alphabet_info, res_info, log):
    """train_flipflop.py:332-461"""
    log.write('* Reading network from {}\n'.format(args.model))
    model_kwargs = {'stride': args.stride, 'winlen': args.winlen, 'insize': 1,
                    'size': args.size, 'alphabet_info': alphabet_info}
    model_metadata = {'reverse': args.reverse, 'standardize': args.standardize}
    network = helpers.load_model(args.model, model_metadata=model_metadata, **model_kwargs)
    log.write('* Network has {} parameters.\n'.format(
        sum(p.nelement() for p in network.parameters())))
    if layers.is_cat_mod_model(network):
        log.write('* Loaded categorical modified base model.\n')
        if not alphabet_info.contains_modified_bases():
            sys.stderr.write('* ERROR: Modified bases model specified, but the input does not '
                             'contain modified bases.')
            sys.exit(1)
    else:
        log.write('* Loaded standard (canonical bases-only) model.\n')
        if alphabet_info.contains_modified_bases():
            sys.stderr.write('* ERROR: Standard model specified, but the input contains '
                             'modified bases.')
            sys.exit(1)
    network = network.to(res_info.device)
    if res_info.is_lead_process:
        log.write('* Dumping initial model\n')
        helpers.save_model(network, args.outdir, 0)
    if res_info.is_multi_gpu:
        # every rank starts from rank 0's weights (the reference reloads checkpoint 0)
        for p in network.parameters():
            torch.distributed.broadcast(p.data, 0)
    network_metadata = training.parse_network_metadata(network)
    stride = helpers.guess_model_stride(network)
    optimiser = torch.optim.AdamW(network.parameters(), lr=args.lr_max, betas=tuple(args.adam),
                                  weight_decay=args.weight_decay, eps=args.eps,
                                  fused=True)   # one multi-tensor kernel per step
    lr_warmup = args.lr_min if args.lr_warmup is None else args.lr_warmup
    adam_beta1, _ = args.adam
    if args.warmup_batches >= args.niteration:
        sys.stderr.write('* Error: --warmup_batches must be < --niteration\n')
        sys.exit(1)
    lr_scheduler = torch.optim.lr_scheduler.OneCycleLR(
        optimiser, args.lr_max, total_steps=args.niteration,
        pct_start=args.warmup_batches / args.niteration, div_factor=args.lr_max / lr_warmup,
        final_div_factor=lr_warmup / args.lr_min,
        cycle_momentum=(args.min_momentum is not None),
        base_momentum=adam_beta1 if args.min_momentum is None else args.min_momentum,
        max_momentum=adam_beta1)
    log.write(('* Learning rate increases from {:.2e} to {:.2e} over {} iterations using cosine '
               'schedule.\n').format(lr_warmup, args.lr_max, args.warmup_batches))
    if args.gradient_clip_num_mads is None:
        log.write('* No gradient clipping\n')
        rolling_mads = None
    else:
        nparams = len([p for p in network.parameters() if p.requires_grad])
        # history and thresholds on the training device: the loop does not have to read the maxima
        # back before it can enqueue the next step
        param_device = next(network.parameters()).device
        rolling_mads = maths.RollingMAD(nparams, args.gradient_clip_num_mads,
                                        device=param_device if param_device.type == 'cuda' else None)
        log.write(('* Gradients will be clipped (by value) at {:3.2f} MADs above the median of '
                   'the last {} gradient maximums.\n').format(rolling_mads.n_mads,
                                                              rolling_mads.window))
    net_info = training.NETWORK_INFO(net=network, net_clone=None, metadata=network_metadata,
                                     stride=stride)
    optim_info = OPTIM_INFO(optimiser=optimiser, lr_warmup=lr_warmup, lr_scheduler=lr_scheduler,
                            rolling_mads=rolling_mads)
    return net_info, optim_info


def compute_filter_params(args, net_info, read_data, log):
    """train_flipflop.py:464-483"""
    sampling_chunk_len = (args.chunk_len_min + args.chunk_len_max) // 2
    sampling_chunk_len = (sampling_chunk_len // net_info.stride) * net_info.stride
    filter_params = chunk_selection.sample_filter_parameters(
        read_data, args.sample_nreads_before_filtering, sampling_chunk_len,
        args.filter_mean_dwell, args.filter_max_dwell, args.filter_min_pass_fraction,
        net_info.stride, args.filter_path_buffer)
    log.write(('* Sampled {} chunks: median(mean_dwell)={:.2f}, mad(mean_dwell)={:.2f}\n').format(
        args.sample_nreads_before_filtering, filter_params.median_meandwell,
        filter_params.mad_meandwell))
    return filter_params


def extract_reporting_data(args, read_data, res_info, alphabet_info, filter_params, net_info,
                           log):
    """Fixed validation batches drawn once (train_flipflop.py:486-529)."""
    if not res_info.is_lead_process or args.reporting_sub_batches <= 0:
        return []
    chunk_len = (args.chunk_len_min + args.chunk_len_max) // 2
    chunk_len = (chunk_len // net_info.stride) * net_info.stride
    batches = list(training.prepare_random_batches(
        read_data, chunk_len, args.min_sub_batch_size, args.reporting_sub_batches,
        alphabet_info, filter_params, net_info, log, select_strands_randomly=False))
    log.write('* Standard loss report: chunk length = {} & sub-batch size = {} for {} '
              'sub-batches.\n'.format(chunk_len, args.min_sub_batch_size,
                                      args.reporting_sub_batches))
    return batches


def log_polka(net_info, train_params, optim_info, time_last, score_smoothed, curr_iter,
              total_samples, total_bases, rejection_dict, log):
    """train_flipflop.py:639-661"""
    time_delta = time.time() - time_last
    log.write(MAIN_LOG_POLKA_TMPLT.format(
        (curr_iter + 1) // DOTROWLENGTH, score_smoothed.value, time_delta,
        total_samples / 1000.0 / time_delta, total_bases / 1000.0 / time_delta,
        optim_info.lr_scheduler.get_last_lr()[0]))
    n_tot = sum(rejection_dict.values())
    n_fail = sum(v for k, v in rejection_dict.items() if k != signal_mapping.Chunk.rej_str_pass)
    if train_params.full_filter_status:
        for k, v in rejection_dict.items():
            log.write(" {}:{} ".format(k, v))
    elif n_tot:
        log.write("  {:.1%} chunks filtered".format(n_fail / n_tot))
    log.write("\n")


def log_validation(net_info, reporting_batch_list, train_params, mod_info, curr_iter, logs):
    """train_flipflop.py:673-684"""
    if not reporting_batch_list:
        return
    t0 = time.time()
    _, rloss, _, total_bases, _ = training.calculate_loss(
        net_info, iter(reporting_batch_list), train_params.sharpen.max,
        mod_info.mod_cat_weights, mod_info.mod_factor.final)
    rloss = float(rloss)
    dt = time.time() - t0
    kbases = total_bases / 1e3
    logs.main.write(MAIN_LOG_VAL_TMPLT.format(curr_iter + 1, rloss, kbases / 1e3, dt, kbases / dt))
    logs.validation.write(VAL_TMPLT.format(curr_iter + 1, rloss))


class TrainLoop:
    """The hot loop, train_flipflop.py:532-627, as an object so that it can be run in
    pieces (`run(n)` continues where the previous call stopped; bench.py times the
    entry point's own loop this way, warm-up excluded).  `train_model` runs it whole."""

    def __init__(self, train_params, net_info, optim_info, res_info, read_data, alphabet_info,
                 filter_params, mod_info, reporting_batch_list, logs):
        self.train_params, self.net_info, self.optim_info = train_params, net_info, optim_info
        self.res_info, self.read_data, self.alphabet_info = res_info, read_data, alphabet_info
        self.filter_params, self.mod_info, self.logs = filter_params, mod_info, logs
        self.reporting_batch_list = reporting_batch_list
        self.step = training.TrainStep(net_info, optim_info.optimiser, optim_info.rolling_mads,
                                       mod_info=mod_info, sub_batches=train_params.sub_batches)
        # reads resident in HBM: batches are assembled by three kernel launches, enqueued one
        # iteration ahead on a side stream (device_batching.BatchPrefetcher), instead of ~5 ms
        # of single-threaded numpy per batch (the step itself takes ~6.5 ms)
        self.prefetcher = None
        device = next(net_info.net.parameters()).device
        if device.type == 'cuda' and not train_params.host_batching:
            store = device_batching.DeviceReadStore(read_data, device)
            self.prefetcher = device_batching.BatchPrefetcher(store, alphabet_info, filter_params,
                                                              net_info, logs.main)
        self.score_smoothed = helpers.WindowedExpSmoother()
        self.total_bases = self.total_samples = 0
        self.samples_seen = 0
        self.rejection_dict = defaultdict(int)
        self.curr_iter = 0
        self.next_shape = None
        self.deferred_log = None
        self.lr_of_iter = {}
        self.iter_times = None
        # Everything built so far (model, optimiser, reads, torch's own module graph) lives for the whole
        # run: take it out of the garbage collector's generations, so that a full collection in the middle
        # of training walks the few objects of a step instead of all of them (pauses of ~50 ms were
        # measured, a whole iteration's worth of device time: `e2e.host_iteration_ms` of bench.py).
        if os.environ.get('TY_GC_FREEZE', '1') != '0':
            import gc
            gc.collect()
            gc.freeze()
        self.time_last = time.time()

    def draw_batch_shape(self):
        """Chunk length and sub-batch size of one iteration (train_flipflop.py:554-562);
        with device batching the iteration's batches are enqueued right away."""
        tp = self.train_params
        batch_chunk_len = (np.random.randint(tp.chunk_len_min, tp.chunk_len_max + 1) //
                           self.net_info.stride) * self.net_info.stride
        sub_batch_size = int(tp.min_sub_batch_size * tp.chunk_len_max / batch_chunk_len + 0.5)
        if self.prefetcher is not None:
            for _ in range(tp.sub_batches):
                self.prefetcher.request(batch_chunk_len, sub_batch_size)
        return batch_chunk_len, sub_batch_size

    def run(self, niter):
        """`niter` more optimiser steps (never beyond train_params.niteration)."""
        tp, net_info, optim_info, mod_info = (self.train_params, self.net_info, self.optim_info,
                                              self.mod_info)
        logs, res_info = self.logs, self.res_info
        last = min(tp.niteration, self.curr_iter + niter)
        if self.next_shape is None and self.curr_iter < tp.niteration:
            self.next_shape = self.draw_batch_shape()
        # With `step.pipelined` (no host-side clipping state) iteration k+1 is enqueued BEFORE the
        # results of iteration k are read back, so the device never drains between optimiser steps;
        # the results of k are processed while k+1 runs.  `in_flight` = the enqueued, unread iteration.
        in_flight = None

        def complete(item):
            pending, curr_iter, batch_chunk_len = item
            (chunk_count, _, chunk_samples, chunk_bases, batch_rejections), fval, grad_maxs = \
                self.step.finish(pending)
            assert np.isfinite(fval), (
                "Error: all costs must be finite, got {}.\n"
                "Try restarting from a checkpoint with a lower learning rate.").format(fval)
            assert np.all(np.isfinite(grad_maxs)), (
                "Error: Gradients not finite.\n"
                "Try restarting from a checkpoint with a lower learning rate.")
            self.samples_seen += chunk_samples
            self.flush_log()
            self.deferred_log = (curr_iter, fval, grad_maxs, self.step.grad_max_threshs,
                                 self.lr_of_iter.pop(curr_iter), batch_chunk_len,
                                 chunk_samples, chunk_bases, batch_rejections)

        while self.curr_iter < last:
            curr_iter = self.curr_iter
            if self.iter_times is not None:          # host time stamp per iteration (bench.py diagnostics)
                self.iter_times.append(time.perf_counter())
            sharpen = float(tp.sharpen.min + (tp.sharpen.max - tp.sharpen.min) *
                            min(1.0, curr_iter / tp.sharpen.niter))
            mod_factor = float(mod_info.mod_factor.start + (
                mod_info.mod_factor.final - mod_info.mod_factor.start) *
                min(1.0, curr_iter / mod_info.mod_factor.niter))
            batch_chunk_len, sub_batch_size = self.next_shape
            if self.prefetcher is not None:
                main_batch_gen = self.prefetcher.batches(tp.sub_batches)
            else:
                main_batch_gen = training.prepare_random_batches(
                    self.read_data, batch_chunk_len, sub_batch_size, tp.sub_batches,
                    self.alphabet_info, self.filter_params, net_info, logs.main)
            self.lr_of_iter[curr_iter] = optim_info.lr_scheduler.get_last_lr()[0]
            pending = self.step.enqueue(main_batch_gen, sharpen, mod_factor)
            # With the step on the stream: the batches of iteration k+1 go onto the batching
            # stream (assembly and its read-back run under this step), and the log lines of
            # iteration k-1 are written -- the device never waits for host bookkeeping.
            if self.prefetcher is not None and curr_iter + 1 < tp.niteration:
                self.next_shape = self.draw_batch_shape()
            if in_flight is not None:
                complete(in_flight)
            in_flight = (pending, curr_iter, batch_chunk_len)
            save_now = (curr_iter + 1) % tp.save_every == 0
            if not self.step.pipelined or save_now or curr_iter + 1 >= last:
                complete(in_flight)
                in_flight = None
            if save_now:
                self.flush_log()
                if res_info.is_lead_process:
                    saved = helpers.save_model(net_info.net, tp.outdir,
                                               (curr_iter + 1) // tp.save_every)
                    logs.main.write("Model saved to {}.\n".format(saved))
                    log_validation(net_info, self.reporting_batch_list, tp, mod_info,
                                   curr_iter, logs)
                self.time_last = time.time()
            optim_info.lr_scheduler.step()
            if self.prefetcher is None and curr_iter + 1 < tp.niteration:
                self.next_shape = self.draw_batch_shape()
            self.curr_iter += 1
        self.flush_log()
        return self.score_smoothed.value

    def flush_log(self):
        """batch.log line, progress dot and the 50-iteration summary of the last finished
        iteration (train_flipflop.py:580-606)."""
        if self.deferred_log is None:
            return
        (curr_iter, fval, grad_maxs, thr, lr, batch_chunk_len, chunk_samples, chunk_bases,
         batch_rejections) = self.deferred_log
        self.deferred_log = None
        logs = self.logs
        if self.res_info.is_lead_process and logs.batch is not None:
            thr_str = 'NaN' if thr is None else ','.join(str(float(t)) for t in thr)
            logs.batch.write(BATCH_TMPLT.format(
                curr_iter + 1, fval, ','.join(map(str, grad_maxs)), thr_str, lr, batch_chunk_len))
        self.total_samples += chunk_samples
        self.total_bases += chunk_bases
        self.score_smoothed.update(fval)
        for k, v in batch_rejections.items():
            self.rejection_dict[k] += v
        logs.main.write('.')
        if (curr_iter + 1) % DOTROWLENGTH == 0:
            log_polka(self.net_info, self.train_params, self.optim_info, self.time_last,
                      self.score_smoothed, curr_iter, self.total_samples, self.total_bases,
                      self.rejection_dict, logs.main)
            self.time_last = time.time()
            self.total_bases = self.total_samples = 0


def train_model(train_params, net_info, optim_info, res_info, read_data, alphabet_info,
                filter_params, mod_info, reporting_batch_list, logs):
    loop = TrainLoop(train_params, net_info, optim_info, res_info, read_data, alphabet_info,
                     filter_params, mod_info, reporting_batch_list, logs)
    if os.environ.get('TY_GRAPHS', '0') == '1' and train_params.sub_batches == 1:
        loop.step.use_graphs(True)
    loop.run(train_params.niteration)
    if res_info.is_lead_process:
        helpers.save_model(net_info.net, train_params.outdir)


def main(args):
    if getattr(args, 'cuda_graphs', False):
        os.environ['TY_GRAPHS'] = '1'
    res_info, logs = parse_init_args(args)
    read_data, alphabet_info, mod_info = load_data(args, logs.main, res_info)
    net_info, optim_info = load_network(args, alphabet_info, res_info, logs.main)
    filter_params = compute_filter_params(args, net_info, read_data, logs.main)
    reporting_batch_list = extract_reporting_data(args, read_data, res_info, alphabet_info,
                                                  filter_params, net_info, logs.main)
    train_params = TRAIN_PARAMS(args.niteration, SHARPEN(*args.sharpen), args.chunk_len_min,
                                args.chunk_len_max, args.min_sub_batch_size, args.sub_batches,
                                args.save_every, args.outdir, args.full_filter_status,
                                args.host_batching)
    train_model(train_params, net_info, optim_info, res_info, read_data, alphabet_info,
                filter_params, mod_info, reporting_batch_list, logs)
    if res_info.is_multi_gpu:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main(get_train_flipflop_parser().parse_args())
