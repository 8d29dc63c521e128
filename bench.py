#!/usr/bin/env python
"""bench.py -- signal-samples/s through the flip-flop CTC/CRF train step.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm, CPU)
    torchrun --nproc-per-node N bench.py --gpus N ...        (N > 1, one rank per GPU)

A step is one optimiser step of bin/train_flipflop.py's loop on one sub-batch
(BASELINE.json configs[1]): mLstm_flipflop, size 256, stride 5, chunk length
T_sig = 4000 samples (nblk = 800), 64 chunks per GPU, 4-base flip-flop (S = 40),
synthetic r9.4.1-like reads, random-init weights:
    zero_grad -> net -> crf_flipflop_loss + flipflop_logpartition / nblk -> backward
    -> gradient all-reduce (N > 1) -> clipping maxima -> AdamW.
Weak scaling: every rank processes its own 64 chunks.

Legs (all in one JSON line printed by rank 0):
  value      K steps with the batches already resident in HBM; CUDA events,
             barrier + synchronize on both sides, max over ranks.
  e2e        K iterations of the ENTRY POINT's own loop (bin/train_flipflop.py:
             TrainLoop.run): candidate draws on the host, H2D, batch assembly on the
             device from reads resident in HBM, the step, D2H of loss + gradient
             maxima + batch counters, logging -- everything the script does per
             iteration is inside the timed region.  `host_signal_leg` keeps the
             round-1 definition (pre-assembled pinned HOST batches through
             training.TrainStep: the whole signal crosses PCIe every step).
  roofline   CRF forward-backward launches (crf_chain_kernel + crf_grad_kernel)
             timed with CUDA events on the launching stream inside the timed
             steps: algorithmic bytes 2*S*4 per (block, chunk) / duration,
             against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's CPU path (stock torch CPU modules + the
             reference's own C loss, oracle/ref_train_step.py) on a bounded
             sample, rank 0 at N = 1 only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_SIG, NCHUNK, SIZE, STRIDE, NTRANS = 4000, 64, 256, 5, 40
METRIC = 'signal_samples_per_sec_flipflop_train_step'
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the loss kernels at this workload, from
# `ncu --set full` captures (profiles/r2_ncu_loss_summary.csv): crf_fused_kernel 33.50 + 67.44 MB,
# logz_chain_kernel 8.24 + 0, logz_post_kernel 11.47 + 0  (round 1, chain + posterior kernel pair
# with the full alpha / beta spill: 356.7 MB)
NCU_TRAFFIC_BYTES = int((33.50 + 67.44 + 8.24 + 0.0 + 11.47 + 0.0) * 1e6)
NCU_TRAFFIC_SOURCE = ('profiles/r2_ncu_loss_summary.csv: dram read+write of crf_fused_kernel + logz_chain_kernel + '
                      'logz_post_kernel, one launch each, config A')
WORKLOAD = 'mLstm_flipflop size256 stride5, T_sig=4000 (nblk=800), 64 chunks/GPU, S=40'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ref-chunks', type=int, default=16,
                    help='chunks per step of the CPU reference sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    # the other BASELINE.json configurations (parity / sweep cases, not the contract line):
    ap.add_argument('--model', default='mLstm_flipflop',
                    choices=['mLstm_flipflop', 'mGru_flipflop', 'mGru_cat_mod_flipflop',
                             'mLstm_cat_mod_flipflop'],
                    help='model definition under models/ (default: the contract workload)')
    ap.add_argument('--tsig', type=int, default=T_SIG, help='chunk length in samples (sweep E)')
    ap.add_argument('--graphs', action='store_true',
                    help='replay forward + loss + backward as a CUDA graph (training.GraphedBody): for the '
                         'launch-bound short-chunk configurations; the chunk length of a bench run is fixed')
    ap.add_argument('--chunks', type=int, default=NCHUNK,
                    help='chunks per GPU and step (BASELINE configs[0]: --model mGru_flipflop --tsig 2000 --chunks 8)')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4),
        }
        # NVML queries contend with kernel launches inside the driver: with a query every 50 ms the
        # launching thread stalled 30-120 ms now and then (one iteration out of ~50; `e2e.host_iteration_ms`,
        # A/B with the sampler off at 4 GPUs).  So: the SM clock every 200 ms, the throttle reasons every
        # fifth sample and once more when the sampler is stopped (still inside the measured part of the run).
        period = float(os.environ.get('TY_BENCH_NVML_PERIOD', '0.2'))

        def reasons():
            try:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
        n = 0
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            except Exception:
                pass
            if n % 5 == 0:
                reasons()
            n += 1
            time.sleep(period)
        reasons()

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the train step on this box's
    host cores, bounded sample of the same workload."""
    if rank != 0:
        return None
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank):
    # rank 0 alone runs this arm, the other ranks have exited
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    os.environ['OMP_NUM_THREADS'] = str(cores)
    os.environ['MKL_NUM_THREADS'] = str(cores)
    import torch
    torch.set_num_threads(cores)
    from oracle import oracle, ref_train_step
    oracle.build()
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    # the contract workload by default; --model / --tsig / --chunks select another BASELINE configuration
    # (configs[0] as written: --model mGru_flipflop --tsig 2000 --chunks 8 --steps 10)
    cell = 'gru' if 'Gru' in args.model else 'lstm'
    stride = 2 if cell == 'gru' else STRIDE
    custom = args.model != 'mLstm_flipflop' or args.tsig != T_SIG or args.chunks != NCHUNK
    ref_chunks = args.chunks if custom else args.ref_chunks
    workload = WORKLOAD if not custom else '%s size%d stride%d, T_sig=%d (nblk=%d), %d chunks, S=40' % (
        args.model.replace('_cat_mod', ''), SIZE, stride, args.tsig, -(-args.tsig // stride), ref_chunks)
    r = ref_train_step.time_reference(cell, args.tsig, ref_chunks, steps, warmup, stride, threads=cores)
    sample = ('%d chunks x T_sig=%d per step, %d timed steps, torch CPU nn.%s + reference C '
              'loss (%s)' % (ref_chunks, args.tsig, steps, 'GRU' if cell == 'gru' else 'LSTM',
                             'oracle/_ref' if r['reference_c'] else 'oracle port'))
    line = {
        'metric': METRIC, 'value': r['samples_per_s'], 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': workload, 'sample': sample},
        'cpu_baseline': {'value': r['samples_per_s'], 'unit': 'samples/s', 'cores': r['threads'],
                         'kind': 'reference' if r['reference_c'] else 'port', 'sample': sample},
        'e2e': {'value': r['samples_per_s'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    return line


class QuietStdout:
    """stdout carries exactly ONE JSON line: while the work runs, file descriptor 1 points
    at stderr, so C-level chatter (NCCL's version banner, library warnings) lands there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def emit(line):
    print(json.dumps(line))
    sys.stdout.flush()


def main():
    with QuietStdout():
        line = run_arm()
    if line is not None:
        emit(line)


def run_arm():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from taiyaki_b200 import _lib, ctc, helpers, signal_mapping, training, chunk_selection
    from taiyaki_b200.alphabet import AlphabetInfo

    assert torch.cuda.is_available(), 'bench.py (our arm) needs a GPU; there is no CPU path'
    _lib.lib()
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log (version banner, NCCL_DEBUG=INFO)
        # goes to stderr unless the caller chose a file
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=device)
    seed = 1 + rank                       # train_flipflop.py:266-268
    np.random.seed(seed)
    torch.manual_seed(seed)

    # ---- model, optimiser (train_flipflop.py:332-429) ----
    global T_SIG, STRIDE, NTRANS, WORKLOAD, NCHUNK
    cat_mod = 'cat_mod' in args.model
    if args.model != 'mLstm_flipflop' or args.tsig != T_SIG or args.chunks != NCHUNK:
        NCHUNK = args.chunks
        args.no_cpu_baseline = True          # the CPU arm times the contract workload only
        T_SIG = args.tsig
        STRIDE = 5 if 'Lstm' in args.model else 2
        NTRANS = 45 if cat_mod else 40
        WORKLOAD = '%s size%d stride%d, T_sig=%d (nblk=%d), %d chunks/GPU, S=%d' % (
            args.model, SIZE, STRIDE, T_SIG, -(-T_SIG // STRIDE), NCHUNK, NTRANS)
    alphabet_info = (AlphabetInfo('ACGTZ', 'ACGTC', ['5mC']) if cat_mod
                     else AlphabetInfo('ACGT', 'ACGT'))
    net = helpers.load_model(os.path.join(ROOT, 'models', args.model + '.py'),
                             model_metadata={'reverse': False, 'standardize': True},
                             stride=STRIDE, winlen=19, insize=1, size=SIZE,
                             alphabet_info=alphabet_info).to(device)
    if world > 1:       # all ranks start from rank 0's weights (checkpoint_00000 in the reference)
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    net_info = training.NETWORK_INFO(net=net, net_clone=None,
                                     metadata=training.parse_network_metadata(net),
                                     stride=STRIDE)
    optimiser = torch.optim.AdamW(net.parameters(), lr=4e-3, betas=(0.9, 0.999),
                                  weight_decay=0.01, eps=1e-6, fused=True)
    mod_info = training.MOD_INFO(np.ones(alphabet_info.nbase, dtype=np.float32), None) \
        if cat_mod else None
    step_fn = training.TrainStep(net_info, optimiser, mod_info=mod_info)
    if args.graphs:
        step_fn.use_graphs(True, min_repeats=1)
    nparam = sum(p.numel() for p in net.parameters() if p.requires_grad)

    # ---- synthetic batches through the reference's batching surface ----
    reads = signal_mapping.synthetic_reads(48, seed=7 + rank, mod_fraction=0.5 if cat_mod else 0.0)
    fp = chunk_selection.sample_filter_parameters(reads, 200, T_SIG, 10.0, 10.0, 0.1, STRIDE, 1.1)
    nbatches = 4
    host_batches = list(training.prepare_random_batches(
        reads, T_SIG, NCHUNK, nbatches, alphabet_info, fp, net_info, None))
    assert all(b[4] == NCHUNK for b in host_batches)
    dev_batches = []
    for indata, seqs, seqlens, mod_cats, nb, rej in host_batches:
        sl_dev = seqlens.to(device)
        ctc.hint_lengths(sl_dev, int(seqlens.max()), int(seqlens.sum()))
        dev_batches.append((indata.to(device), seqs.to(device), sl_dev,
                            None if mod_cats is None else mod_cats.to(device), nb, rej))
    h2d = int(sum(b[0].numel() * 4 + b[1].numel() * 8 + b[2].numel() * 8
                  for b in host_batches) / nbatches)

    def run(batches, steps, read_back):
        """read_back: every step's loss and gradient maxima are read on the host (one D2H per step);
        as in the entry point's loop, step k+1 is enqueued before the results of step k are read."""
        out = in_flight = None
        for i in range(steps):
            batch = iter([batches[i % len(batches)]])
            if not read_back:
                out = step_fn(batch, sharpen=1.0, mod_factor=1.0, read_back=False)
                continue
            pending = step_fn.enqueue(batch, 1.0, 1.0)
            if in_flight is not None:
                out = step_fn.finish(in_flight)
            in_flight = pending
        if in_flight is not None:
            out = step_fn.finish(in_flight)
        return out

    def timed(batches, steps, read_back):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = run(batches, steps, read_back)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    K, W = args.steps, max(args.warmup, 3)
    run(dev_batches, W, False)
    run(host_batches, 2, True)
    assert step_fn.flat.check_views(), 'gradient views detached from the flat buffer'

    sampler = ClockSampler(local_rank)
    if os.environ.get('TY_BENCH_NO_NVML', '0') == '1':      # diagnostics: is the sampler itself the disturbance?
        sampler.nv = None
    sampler.start()
    launches0 = _lib.LAUNCHES
    _lib.PROFILE = {}
    ms_dev, _ = timed(dev_batches, K, False)
    launches = _lib.LAUNCHES - launches0
    prof = _lib.PROFILE
    _lib.PROFILE = None
    ms_host, out = timed(host_batches, K, True)
    loss = out[1]
    assert np.isfinite(loss), 'non-finite loss'

    # ---- e2e: the entry point's own loop (bin/train_flipflop.py: TrainLoop.run) ----
    # what `train_flipflop.py model.py reads.hdf5` executes per iteration: draw the chunk
    # length and the candidate windows on the host, H2D of the candidates, batch assembly
    # on the device from the reads resident in HBM (uploaded once), the train step, the
    # D2H of loss + gradient maxima + batch counters, the batch log line.
    import importlib.util
    spec = importlib.util.spec_from_file_location('train_flipflop_entry',
                                                  os.path.join(ROOT, 'bin', 'train_flipflop.py'))
    tf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tf)

    class NullLog:
        def write(self, msg):
            pass

    class ConstantLR:
        def get_last_lr(self):
            return [4e-3]

        def step(self):
            pass
    train_params = tf.TRAIN_PARAMS(
        niteration=10 ** 9, sharpen=tf.SHARPEN(1.0, 1.0, 1), chunk_len_min=T_SIG,
        chunk_len_max=T_SIG, min_sub_batch_size=NCHUNK, sub_batches=1, save_every=10 ** 9,
        outdir=None, full_filter_status=False, host_batching=False)
    loop = tf.TrainLoop(
        train_params, net_info, tf.OPTIM_INFO(optimiser, 4e-3, ConstantLR(), None),
        tf.RESOURCE_INFO(world > 1, rank == 0, device), reads, alphabet_info, fp,
        training.MOD_INFO(np.ones(alphabet_info.nbase, dtype=np.float32), tf.MOD_FACTOR(1.0, 1.0, 1)),
        [], tf.LOGS(main=NullLog()))
    if args.graphs:
        loop.step.use_graphs(True, min_repeats=1)
    loop.run(W)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc_pauses, gc_t0 = [], [0.0]

    def gc_cb(phase, info):
        if phase == 'start':
            gc_t0[0] = time.perf_counter()
        else:
            gc_pauses.append((time.perf_counter() - gc_t0[0]) * 1e3)
    gc.callbacks.append(gc_cb)
    seen0, wall0 = loop.samples_seen, time.perf_counter()
    loop.iter_times = []
    ev_a.record()
    loop.run(K)
    ev_b.record()
    torch.cuda.synchronize()
    wall_entry = time.perf_counter() - wall0
    iter_ms = np.diff(np.array(loop.iter_times + [time.perf_counter()])) * 1e3
    loop.iter_times = None
    gc.callbacks.remove(gc_cb)
    ms_entry = torch.tensor([max(ev_a.elapsed_time(ev_b), wall_entry * 1e3)], device=device)
    entry_samples = torch.tensor([float(loop.samples_seen - seen0)], device=device)
    if world > 1:
        dist.all_reduce(ms_entry, op=dist.ReduceOp.MAX)
        dist.all_reduce(entry_samples, op=dist.ReduceOp.SUM)
    ms_entry, entry_samples = float(ms_entry), float(entry_samples)
    gc_max = torch.tensor([max(gc_pauses) if gc_pauses else 0.0], device=device)
    if world > 1:
        dist.all_reduce(gc_max, op=dist.ReduceOp.MAX)
    gc_max = float(gc_max)
    store = loop.prefetcher.store
    attempts = max(NCHUNK, int(NCHUNK / fp.filter_min_pass_fraction))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    samples_per_step = T_SIG * NCHUNK * world
    value = samples_per_step * K / (ms_dev * 1e-3)
    e2e = entry_samples / (ms_entry * 1e-3)
    e2e_host = samples_per_step * K / (ms_host * 1e-3)

    # ---- roofline of the CRF forward-backward launches ----
    peak, peak_src = measured_peak()
    nblk = -(-T_SIG // STRIDE)
    alg_bytes = 2 * NTRANS * 4 * nblk * NCHUNK
    key = 'loss_fwd_bwd' if 'loss_fwd_bwd' in prof else 'crf_fwd_bwd'
    crf_ms = [a.elapsed_time(b) for a, b in prof.get(key, [])]
    rnn_ms = [a.elapsed_time(b) for a, b in prof.get('rnn_fwd', [])]
    rnnb_ms = [a.elapsed_time(b) for a, b in prof.get('rnn_bwd', [])]
    crf_avg = float(np.mean(crf_ms)) if crf_ms else float('nan')
    achieved = alg_bytes / (crf_avg * 1e-3) / 1e9
    fused_crf = _lib.lib().ty_crf_last_path() == 2
    crf_kernels = ('crf_fused_kernel (label-constrained chains with the posterior fused in)' if fused_crf
                   else 'crf_chain_kernel, crf_post_kernel')
    kname = ('ty_flipflop_train_loss: %s || logz_chain_kernel, logz_post_kernel (CRF loss + logZ/nblk, '
             'forward-backward and gradient)' % crf_kernels if key == 'loss_fwd_bwd'
             else crf_kernels + ' (CRF fwd-bwd)')
    roofline = {'bound': 'hbm', 'kernel': kname,
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': NCU_TRAFFIC_BYTES if (key == 'loss_fwd_bwd' and fused_crf and T_SIG == 4000
                                                 and NTRANS == 40) else None,
                'traffic_source': NCU_TRAFFIC_SOURCE,
                'peak_source': peak_src, 'algorithmic_bytes': alg_bytes,
                'avg_launch_ms': crf_avg, 'launches_timed': len(crf_ms),
                'share_of_step': crf_avg / (ms_dev / K),
                'note': 'latency-bound sequential DP (nblk dependent steps); see DESIGN.md 4.1'}
    # ---- batch assembly (outside the timed legs): host numpy path vs device kernels ----
    batching = None
    if rank == 0:
        from taiyaki_b200 import device_batching
        t0 = time.time()
        nb = len(list(training.prepare_random_batches(reads, T_SIG, NCHUNK, 5, alphabet_info, fp,
                                                      net_info, None)))
        host_ms = (time.time() - t0) * 1e3 / nb
        store = loop.prefetcher.store
        list(device_batching.prepare_random_batches(store, T_SIG, NCHUNK, 2, alphabet_info, fp,
                                                    net_info, None))
        torch.cuda.synchronize()
        t0 = time.time()
        nb = len(list(device_batching.prepare_random_batches(store, T_SIG, NCHUNK, 10,
                                                             alphabet_info, fp, net_info, None)))
        torch.cuda.synchronize()
        batching = {'host_ms_per_batch': host_ms,
                    'device_ms_per_batch': (time.time() - t0) * 1e3 / nb,
                    'note': 'one batch, launch + wait, NOT pipelined (chunk sampling, filters, '
                            'stacking, flip-flop coding); inside the e2e leg the device path runs one '
                            'iteration ahead on a side stream'}
    # the recurrent kernels are the only tensor-core work of the path: achieved
    # TFLOP/s of the per-step products against the sustained bf16 peak, for context
    rnn_tensor = None
    if rnn_ms and rnnb_ms and args.model == 'mLstm_flipflop':
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops_sustained'])
            tsrc = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
        except Exception:
            tpeak, tsrc = 1377.0, 'fallback (SURVEY 8d)'
        flops = 2.0 * nblk * NCHUNK * 4 * SIZE * SIZE          # W_hh h per layer and direction
        rnn_tensor = {
            'bound': 'tensor (latency-bound recurrence: nblk dependent steps, DESIGN.md 4.5)',
            'fwd_tflops': flops / (np.mean(rnn_ms) * 1e-3) / 1e12,
            'bwd_tflops': flops / (np.mean(rnnb_ms) * 1e-3) / 1e12,
            'peak_tflops': tpeak, 'peak_source': tsrc,
            'frac_fwd': flops / (np.mean(rnn_ms) * 1e-3) / 1e12 / tpeak}
    extra = {'batching': batching, 'rnn_tensor': rnn_tensor,
             'rnn_fwd_kernel_ms_avg': float(np.mean(rnn_ms)) if rnn_ms else None,
             'rnn_bwd_kernel_ms_avg': float(np.mean(rnnb_ms)) if rnnb_ms else None,
             'rnn_layers': 5, 'trainable_params': nparam, 'loss': loss}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle, ref_train_step
            oracle.build()
            r = ref_train_step.time_reference('lstm', T_SIG, args.ref_chunks, 2, 1, STRIDE,
                                              threads=len(os.sched_getaffinity(0)))
            cpu_baseline = {
                'value': r['samples_per_s'], 'unit': 'samples/s', 'cores': r['threads'],
                'kind': 'reference' if r['reference_c'] else 'port',
                'sample': '%d chunks x T_sig=%d, 2 timed steps after 1 warm-up; stock torch CPU '
                          'nn.LSTM stack + the reference C loss via oracle/_ref'
                          % (args.ref_chunks, T_SIG)}
        except Exception as e:     # the baseline is informational; never fail the bench on it
            cpu_baseline = {'value': None, 'error': str(e)[:200]}
        # the reference's GPU path for the partition function (its CuPy RawKernels compiled with nvcc by
        # oracle/build_cupy_ref.py) next to csrc/logz.cu on the same scores, for context
        try:
            from oracle import oracle
            from taiyaki_b200 import layers
            if oracle.libcupy_ref() is not None:
                sc = (5.0 * torch.tanh(torch.randn(nblk, NCHUNK, 40, device=device))).contiguous()

                def gpu_ms(fn, reps=5):
                    fn()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(reps):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        fn()
                        b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                    return float(np.median(ts))

                def ours():
                    x = sc.detach().requires_grad_(True)
                    layers.flipflop_logpartition(x).sum().backward()
                extra['reference_gpu_path'] = {
                    'what': 'logZ + gradient of [%d, %d, 40] scores: the reference CuPy RawKernels '
                            '(cupy_extensions/flipflop.py:10-296) under nvcc vs csrc/logz.cu' % (nblk, NCHUNK),
                    'reference_ms': gpu_ms(lambda: oracle.cupy_ref_logz(sc)), 'ours_ms': gpu_ms(ours)}
        except Exception as e:
            extra['reference_gpu_path'] = {'error': str(e)[:200]}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms_dev / K, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16 tensor-core products (fp32 accumulate) / f32 state, gates and loss',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'chunks_per_gpu': NCHUNK, 'global_chunks': NCHUNK * world,
                   'parallelism': 'dp%d' % world, 'cuda_graphs': bool(args.graphs),
                   'l2': 'per-step working set (activations + reserve, >2 GB) exceeds the 126 MB L2'},
        'e2e': {'value': e2e, 'unit': 'samples/s',
                'path': "bin/train_flipflop.py TrainLoop.run (what the entry point executes): host "
                        "draws chunk length + candidate windows -> H2D -> batch assembly on the device "
                        "from reads resident in HBM, one iteration ahead on a side stream -> train step "
                        "-> D2H of loss, gradient maxima and batch counters -> batch log",
                'h2d_bytes_per_step': 2 * attempts * 4,
                'd2h_bytes_per_step': 4 * (1 + len(step_fn.flat.params)) + (2 * NCHUNK + 1) * 8 + 64,
                'reads_resident_bytes': int(store.dacs.numel() * 2 + store.r2s.numel() * 4 +
                                            store.ref.numel() * 2),
                'ms_per_step': ms_entry / K, 'fraction_of_value': e2e / value,
                'host_iteration_ms': {'median': float(np.median(iter_ms)), 'max': float(iter_ms.max()),
                                      'argmax': int(iter_ms.argmax()),
                                      'gc_collections': len(gc_pauses),
                                      'gc_pause_ms_max_over_ranks': gc_max},
                'timed_as': 'max(CUDA events, host wall clock) around K iterations, max over ranks',
                'host_signal_leg': {
                    'value': e2e_host, 'unit': 'samples/s', 'ms_per_step': ms_host / K,
                    'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4 * (1 + len(step_fn.flat.params)),
                    'path': 'training.TrainStep on pre-assembled pinned HOST batches: the whole signal '
                            'tensor and the labels cross PCIe every step and loss + gradient maxima are read back '
                            'every step, one step in flight (round-1 definition of e2e)'}},
        'gpu_launches': launches,
        'clocks': sampler.summary(),
        'roofline': roofline,
        'cpu_baseline': cpu_baseline,
        'detail': extra,
    }
    if world > 1:
        dist.destroy_process_group()
    return line



if __name__ == '__main__':
    main()
