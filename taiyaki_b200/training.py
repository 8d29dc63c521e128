"""One optimiser step of flip-flop training -- the functions of
bin/train_flipflop.py that sit on the hot path, kept under their own names:

  prepare_random_batches   train_flipflop.py:78-142   (host batching surface)
  calculate_loss           train_flipflop.py:145-198  (net -> CRF loss + logZ/nblk -> backward)
  apply_clipping           train_flipflop.py:201-212
  FlatGradients            replaces DistributedDataParallel's bucketed all-reduce
                           (train_flipflop.py:395-397) by ONE NCCL all-reduce of
                           a flat fp32 gradient buffer per optimiser step.

Differences from the reference are confined to where the work happens: scores
stay on the device through the loss, and the per-tensor `float(max|grad|)` host
round trips of apply_clipping become one device reduction and one copy.
"""
import os
from collections import defaultdict, namedtuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, chunk_selection, ctc, flipflopfings, layers

NETWORK_METADATA = namedtuple('NETWORK_METADATA', (
    'reverse', 'standardize', 'is_cat_mod', 'can_mods_offsets', 'can_labels', 'mod_labels'))
NETWORK_METADATA.__new__.__defaults__ = (None, None, None)
NETWORK_INFO = namedtuple('NETWORK_INFO', ('net', 'net_clone', 'metadata', 'stride'))
MOD_INFO = namedtuple('MOD_INFO', ('mod_cat_weights', 'mod_factor'))

#: use ctc.flipflop_train_loss (one fused operator) inside flipflop_loss
FUSED_LOSS = True
#: run the recurrent layers' weight-gradient GEMMs on a side stream during backward
#: (TY_DEFER_WGRAD=0 in the environment turns it off, for A/B timing)
DEFER_WEIGHT_GRADS = os.environ.get('TY_DEFER_WGRAD', '1') != '0'


def parse_network_metadata(network):
    """train_flipflop.py:67-75"""
    if layers.is_cat_mod_model(network):
        last = network.sublayers[-1]
        return NETWORK_METADATA(network.metadata['reverse'], network.metadata['standardize'],
                                True, last.can_mods_offsets, last.can_labels, last.mod_labels)
    return NETWORK_METADATA(network.metadata['reverse'], network.metadata['standardize'], False)


def prepare_random_batches(read_data, batch_chunk_len, sub_batch_size, target_sub_batches,
                           alphabet_info, filter_params, net_info, log,
                           select_strands_randomly=True, first_strand_index=0, pin=True):
    """Generator of (indata [T,N,1] pinned fp32, seqs, seqlens, mod_cats,
    sub_batch_size, rejections) -- train_flipflop.py:78-142."""
    total_sub_batches = 0
    revop = np.flip if net_info.metadata.reverse else np.array
    while total_sub_batches < target_sub_batches:
        chunk_batch, batch_rejections = chunk_selection.sample_chunks(
            read_data, sub_batch_size, batch_chunk_len, filter_params,
            standardize=net_info.metadata.standardize,
            select_strands_randomly=select_strands_randomly,
            first_strand_index=first_strand_index)
        first_strand_index += sum(batch_rejections.values())
        if len(chunk_batch) < sub_batch_size and log is not None:
            log.write(('* Warning: only {} chunks passed filters (asked for {}).\n').format(
                len(chunk_batch), sub_batch_size))
        if not all(chunk.seq_len > 0.0 for chunk in chunk_batch):
            raise Exception('Error: zero length sequence')
        # [T, N] float32, filled column by column (torch.tensor() of a transposed float64
        # stack, as the reference builds it, costs ~50 ms for 64 x 4000 samples)
        stacked_current = np.empty((len(chunk_batch[0].current), len(chunk_batch)), dtype=np.float32)
        for i, chunk in enumerate(chunk_batch):
            stacked_current[:, i] = revop(chunk.current)
        indata = torch.from_numpy(stacked_current).unsqueeze(2)
        if pin and torch.cuda.is_available():
            indata = indata.pin_memory()
        seqlens = [len(chunk.sequence) for chunk in chunk_batch]
        labels = np.concatenate([revop(chunk.sequence) for chunk in chunk_batch]).astype(np.int64)
        mod_cats = None
        if net_info.metadata.is_cat_mod:
            mod_cats = torch.from_numpy(np.ascontiguousarray(
                net_info.metadata.mod_labels[labels]).astype(np.int64))
            labels = np.ascontiguousarray(net_info.metadata.can_labels[labels]).astype(np.int64)
        # flip-flop coding of all chunks at once: runs cannot cross a chunk boundary
        starts = np.cumsum([0] + seqlens[:-1])
        seqs = torch.from_numpy(flipflopfings.flipflop_code_batch(
            labels, starts, alphabet_info.ncan_base))
        seqlens = torch.tensor(seqlens, dtype=torch.long, device='cpu')
        total_sub_batches += 1
        yield indata, seqs, seqlens, mod_cats, len(chunk_batch), batch_rejections


def flipflop_loss(outputs, seqs, seqlens, sharpen, mod_cats=None, can_mods_offsets=None,
                  mod_cat_weights=None):
    """Per-chunk loss vector: CRF cost + logZ / nblk (train_flipflop.py:163-176).
    FUSED_LOSS selects the single fused operator (same numbers, chains overlapped,
    one gradient write) or the reference's two separate operators."""
    if FUSED_LOSS:
        return ctc.flipflop_train_loss(outputs, seqs, seqlens, sharpen, mod_cats,
                                       can_mods_offsets, mod_cat_weights)
    nblk = float(outputs.shape[0])
    ntrans = outputs.shape[2]
    if mod_cats is not None:
        lossvector = ctc.cat_mod_flipflop_loss(outputs, seqs, seqlens, mod_cats,
                                               can_mods_offsets, mod_cat_weights, sharpen)
        ntrans -= int(can_mods_offsets[-1])
    else:
        lossvector = ctc.crf_flipflop_loss(outputs, seqs, seqlens, sharpen)
    return lossvector + layers.flipflop_logpartition(outputs[:, :, :ntrans]) / nblk


def calculate_loss(net_info, batch_gen, sharpen, mod_cat_weights=None, mod_factor=None,
                   calc_grads=False, device=None):
    """train_flipflop.py:145-198.  Returns (chunk_count, mean loss as a DEVICE
    tensor, samples, bases, rejection dict); the caller decides when to read
    the loss back (one host sync per optimiser step instead of one per
    sub-batch)."""
    can_mods_offsets = net_info.metadata.can_mods_offsets
    total_chunk_count = total_samples = total_bases = n_subbatches = 0
    total_fval = None
    rejection_dict = defaultdict(int)
    device = device if device is not None else next(net_info.net.parameters()).device
    for (indata, seqs, seqlens, mod_cats, sub_batch_size, batch_rejections) in batch_gen:
        n_subbatches += 1
        for k, v in batch_rejections.items():
            rejection_dict[k] += v
        total_chunk_count += sub_batch_size
        with torch.set_grad_enabled(calc_grads):
            outputs = net_info.net(indata.to(device, non_blocking=True))
            if net_info.metadata.is_cat_mod:
                lossvector = flipflop_loss(outputs, seqs, seqlens, sharpen, mod_cats,
                                           can_mods_offsets, mod_cat_weights * mod_factor)
            else:
                lossvector = flipflop_loss(outputs, seqs, seqlens, sharpen)
            loss = lossvector.mean()
        if calc_grads:
            # weight-gradient GEMMs of the recurrent layers overlap the next layer's
            # backward recurrence on a side stream (layers.DEFER_WEIGHT_GRADS)
            prev = layers.DEFER_WEIGHT_GRADS
            layers.DEFER_WEIGHT_GRADS = DEFER_WEIGHT_GRADS and loss.is_cuda
            try:
                loss.backward()
            finally:
                layers.DEFER_WEIGHT_GRADS = prev
                layers.flush_weight_grads()
        total_fval = loss.detach() if total_fval is None else total_fval + loss.detach()
        total_samples += int(indata.nelement())
        total_bases += ctc._max_len(seqlens)[1]     # host value (or a registered hint)
    if calc_grads and n_subbatches > 1:
        for p in net_info.net.parameters():
            if p.grad is not None:
                p.grad /= n_subbatches
    return (total_chunk_count, total_fval / n_subbatches, total_samples, total_bases,
            rejection_dict)


class FlatGradients:
    """All trainable gradients as views of one flat fp32 buffer.

    `zero()` clears it, `all_reduce()` averages it across ranks with a single
    collective (NCCL over NVLink on GPUs, gloo in the CPU tests).  Data
    parallelism is the reference's only strategy (train_flipflop.py:255-268,
    :395-397): chunks are independent and only the weight gradient is shared.
    """

    def __init__(self, parameters, process_group=None):
        self.params = [p for p in parameters if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1

    def zero(self):
        self.flat.zero_()

    def check_views(self):
        """Autograd accumulates in place, so the views must still alias `flat`."""
        base = self.flat.data_ptr()
        return all(p.grad is not None and
                   base <= p.grad.data_ptr() < base + self.flat.numel() * 4
                   for p in self.params)

    # ---- overlapped reduction (one sub-batch per step, NCCL) ----------------------
    # Backward runs from the last layer to the first, so the gradients of the upper layers
    # are final long before backward() returns.  `begin_overlap` arms a hook that the
    # recurrent layers call at the start of their backward (layers.GRADS_FINAL_ABOVE_HOOK):
    # the slice of the flat buffer belonging to the layers above is then all-reduced on the
    # side stream (behind the weight-gradient GEMMs that write into it), under the backward
    # recurrences of the layers below.  `all_reduce` reduces what is left and joins.
    def begin_overlap(self):
        if self.world <= 1 or not self.flat.is_cuda or dist.get_backend(self.group) != 'nccl':
            return False
        if not hasattr(self, '_end_offset'):
            self._end_offset, off = {}, 0
            for p in self.params:
                off += p.numel()
                self._end_offset[id(p)] = off
        self._reduced_from = self.flat.numel()
        self._works = []
        layers.GRADS_FINAL_ABOVE_HOOK = self._grads_final_above
        return True

    def _grads_final_above(self, last_param, side):
        start = self._end_offset.get(id(last_param))
        if start is None or start >= self._reduced_from:
            return
        main = torch.cuda.current_stream(self.flat.device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            self._works.append(dist.all_reduce(self.flat[start:self._reduced_from],
                                               op=dist.ReduceOp.AVG, group=self.group,
                                               async_op=True))
        self._reduced_from = start

    def all_reduce(self):
        if self.world <= 1:
            return
        if layers.GRADS_FINAL_ABOVE_HOOK is not None:       # overlapped mode: reduce the rest, join
            layers.GRADS_FINAL_ABOVE_HOOK = None
            if self._reduced_from > 0:
                dist.all_reduce(self.flat[:self._reduced_from], op=dist.ReduceOp.AVG, group=self.group)
            for w in self._works:
                w.wait()                                    # current stream waits for the NCCL stream
            self._works = []
            return
        if self.flat.is_cuda and dist.get_backend(self.group) == 'nccl':
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)

    def grad_maxs(self):
        """max|grad| per parameter tensor as ONE device tensor."""
        return torch.stack(torch._foreach_norm([p.grad for p in self.params], float('inf')))


def apply_clipping(net_info, grad_max_threshs, flat=None):
    """Clip each parameter tensor by value at its threshold
    (train_flipflop.py:201-212).  Returns the per-tensor maxima as a device
    tensor (read back together with the loss)."""
    parameters = flat.params if flat is not None else \
        [p for p in net_info.net.parameters() if p.requires_grad]
    grad_maxs = torch.stack(torch._foreach_norm([p.grad for p in parameters], float('inf')))
    if grad_max_threshs is not None:
        if torch.is_tensor(grad_max_threshs):     # device-resident thresholds (maths.RollingMAD(device=...))
            thr = grad_max_threshs.to(device=grad_maxs.device, dtype=torch.float32)
        else:
            thr = torch.as_tensor(np.asarray(grad_max_threshs, dtype=np.float32),
                                  device=grad_maxs.device)
        for p, t in zip(parameters, thr):
            # clamp to [-t, t] is a no-op where max|grad| <= t
            torch.minimum(torch.maximum(p.grad, -t, out=p.grad), t, out=p.grad)
    return grad_maxs


class GraphedBody:
    """zero_grad -> forward -> loss -> backward of ONE sub-batch as a CUDA graph, replayed with the
    batch copied into static buffers.  For launch-bound steps (short chunks: ~110 launches of a few
    microseconds of work each, the host cannot keep ahead) when the chunk length is FIXED
    (`--chunk_len_min == --chunk_len_max`; the default entry point draws a new length every iteration,
    train_flipflop.py:554-562, and then every step has another shape -- `TrainStep` falls back to
    eager launches for any shape it has not captured `min_repeats` times).

    What makes the body capturable: everything in it is stream-ordered on the capturing stream (the
    loss's and the weight-gradient GEMMs' side streams fork from and join it by events), gradients
    accumulate into the static flat buffer, and the only data-dependent sizes -- the number of labels
    and the longest chunk -- enter the kernels as CAPACITIES: label tensors are padded to N * Lcap
    entries (valid labels, ignored beyond each chunk's length) and the loss kernels are launched for
    Lcap = the longest chunk rounded up to 64 positions, one graph per (T, N, Lcap, sharpen, ...)."""

    LCAP_STEP = 64
    #: a captured graph owns the activations of a whole step (gigabytes at training sizes): shapes
    #: beyond this many stay eager, so a run with randomly drawn chunk lengths cannot exhaust memory
    MAX_GRAPHS = 8

    def __init__(self, net_info, flat, min_repeats=3):
        self.net_info, self.flat, self.min_repeats = net_info, flat, min_repeats
        self.entries, self.seen = {}, defaultdict(int)
        self.replays = 0

    def _key(self, batch, sharpen, mcw_scaled):
        indata, seqs, seqlens, mod_cats = batch[:4]
        max_len, total = ctc._max_len(seqlens)
        lcap = -(-max(1, max_len) // self.LCAP_STEP) * self.LCAP_STEP
        mk = None if mcw_scaled is None else tuple(np.asarray(mcw_scaled, dtype=np.float32).tolist())
        return (tuple(indata.shape), lcap, float(sharpen), mk, mod_cats is not None), total

    def _body(self, e, sharpen, mcw_scaled):
        can_mods_offsets = self.net_info.metadata.can_mods_offsets
        self.flat.zero()
        outputs = self.net_info.net(e['indata'])
        if self.net_info.metadata.is_cat_mod:
            lossvector = flipflop_loss(outputs, e['seqs'], e['seqlens'], sharpen, e['mod_cats'],
                                       can_mods_offsets, mcw_scaled)
        else:
            lossvector = flipflop_loss(outputs, e['seqs'], e['seqlens'], sharpen)
        loss = lossvector.mean()
        prev = layers.DEFER_WEIGHT_GRADS
        layers.DEFER_WEIGHT_GRADS = DEFER_WEIGHT_GRADS
        try:
            loss.backward()
        finally:
            layers.DEFER_WEIGHT_GRADS = prev
            layers.flush_weight_grads()
        return loss.detach()

    def _capture(self, key, batch, sharpen, mcw_scaled):
        (shape, lcap, _, _, has_mod) = key
        indata, seqs, seqlens, mod_cats = batch[:4]
        dev = indata.device
        N = shape[1]
        e = {'indata': torch.empty(shape, dtype=torch.float32, device=dev),
             'seqs': torch.zeros(N * lcap, dtype=torch.int64, device=dev),
             'seqlens': torch.ones(N, dtype=torch.int64, device=dev),
             'mod_cats': torch.zeros(N * lcap, dtype=torch.int64, device=dev) if has_mod else None}
        ctc.hint_lengths(e['seqlens'], lcap, N * lcap)       # capacities, not this batch's values
        self._load(e, batch)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                        # warm-up outside capture (allocator, lazy init)
            self._body(e, sharpen, mcw_scaled)
        torch.cuda.current_stream(dev).wait_stream(side)
        launches0 = _lib.LAUNCHES
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            e['loss'] = self._body(e, sharpen, mcw_scaled)
        e['graph'], e['launches'] = graph, _lib.LAUNCHES - launches0
        return e

    @staticmethod
    def _load(e, batch):
        indata, seqs, seqlens, mod_cats = batch[:4]
        e['indata'].copy_(indata, non_blocking=True)
        n = seqs.numel()
        e['seqs'][:n].copy_(seqs, non_blocking=True)
        e['seqlens'].copy_(seqlens, non_blocking=True)
        if e['mod_cats'] is not None:
            e['mod_cats'][:n].copy_(mod_cats, non_blocking=True)

    def run(self, batch, sharpen, mcw_scaled):
        """Loss (device scalar) of the batch with its gradients in the flat buffer, or None when this
        shape is still run eagerly."""
        indata, seqs, seqlens = batch[:3]
        if not (indata.is_cuda and seqs.is_cuda and seqlens.is_cuda):
            return None
        key, total = self._key(batch, sharpen, mcw_scaled)
        e = self.entries.get(key)
        if e is None:
            if len(self.entries) >= self.MAX_GRAPHS:
                return None
            if len(self.seen) > 4096:      # (sharpening / mod-factor ramps give every step a new key)
                self.seen.clear()
            self.seen[key] += 1
            if self.seen[key] < self.min_repeats:
                return None
            try:
                e = self.entries[key] = self._capture(key, batch, sharpen, mcw_scaled)
            except Exception as err:      # a step that cannot be captured stays eager, for good
                import warnings
                warnings.warn('taiyaki_b200: CUDA-graph capture of the train step failed (%s); '
                              'continuing with eager launches' % str(err).splitlines()[0][:200])
                self.MAX_GRAPHS = 0
                torch.cuda.synchronize()
                return None
        self._load(e, batch)
        e['graph'].replay()
        _lib.count_launches(e['launches'])
        self.replays += 1
        return e['loss']


class TrainStep:
    """zero_grad -> calculate_loss(calc_grads) -> all-reduce -> clipping ->
    AdamW step (train_flipflop.py:570-578) as one callable."""

    def __init__(self, net_info, optimiser, rolling_mads=None, lr_scheduler=None,
                 mod_info=None, sub_batches=1):
        self.sub_batches = sub_batches      # > 1: gradients accumulate over sub-batches, reduce at the end
        self.net_info = net_info
        self.optimiser = optimiser
        self.rolling_mads = rolling_mads
        self.lr_scheduler = lr_scheduler
        self.mod_info = mod_info
        self.flat = FlatGradients(net_info.net.parameters())
        self.grad_max_threshs = None        # host copy (what batch.log prints)
        self._thr_dev = None                # device-resident thresholds of the next step's clipping
        self._host_bufs = [None, None]      # two steps can be in flight (see `pipelined`)
        self._nstep = 0
        #: CUDA-graph replay of forward + loss + backward for repeated shapes (see GraphedBody);
        #: off by default, `use_graphs()` / TY_GRAPHS=1 turn it on
        self.graphed = GraphedBody(net_info, self.flat) if os.environ.get('TY_GRAPHS', '0') == '1' else None

    def use_graphs(self, on=True, min_repeats=3):
        self.graphed = GraphedBody(self.net_info, self.flat, min_repeats) if on else None
        return self

    @property
    def pipelined(self):
        """True when step k+1 may be enqueued before the results of step k have been read back:
        the clipping thresholds either do not exist or live on the device."""
        return self.rolling_mads is None or getattr(self.rolling_mads, 'device', None) is not None

    def enqueue(self, batch_gen, sharpen=1.0, mod_factor=1.0):
        """Put one optimiser step on the stream without waiting for it; `finish` reads its
        results back.  Splitting the two lets the caller enqueue the next batch's assembly
        (and do its host bookkeeping) while the device is busy with this step."""
        mcw = None if self.mod_info is None else self.mod_info.mod_cat_weights
        res = None
        if self.graphed is not None and self.sub_batches == 1:
            batches = list(batch_gen)
            if len(batches) == 1:
                b = batches[0]
                scaled = mcw * mod_factor if (mcw is not None and self.net_info.metadata.is_cat_mod) else None
                loss = self.graphed.run(b, sharpen, scaled)
                if loss is not None:       # gradients are in the flat buffer (zeroed inside the graph)
                    rej = defaultdict(int)
                    for k, v in b[5].items():
                        rej[k] += v
                    res = (b[4], loss, int(b[0].nelement()), ctc._max_len(b[2])[1], rej)
            batch_gen = iter(batches)
        if res is None:
            self.flat.zero()
            if self.sub_batches == 1 and DEFER_WEIGHT_GRADS:
                self.flat.begin_overlap()
            try:
                res = calculate_loss(self.net_info, batch_gen, sharpen, mcw, mod_factor,
                                     calc_grads=True)
            finally:
                if layers.GRADS_FINAL_ABOVE_HOOK is not None and self.flat.world <= 1:
                    layers.GRADS_FINAL_ABOVE_HOOK = None
        self.flat.all_reduce()
        device_thr = self.rolling_mads is not None and getattr(self.rolling_mads, 'device', None) is not None
        grad_maxs = apply_clipping(self.net_info, self._thr_dev if device_thr else self.grad_max_threshs,
                                   self.flat)
        self.optimiser.step()
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        if device_thr:      # thresholds of the NEXT step, computed where the maxima are (train_flipflop.py:577-578)
            self._thr_dev = self.rolling_mads.update(grad_maxs)
        # one device->host copy per optimiser step: loss, gradient maxima, (device thresholds for the
        # log,) and the label-range flag of the loss operators (ctc.pyx:133-134)
        flag = ctc.pending_flags(res[1].device) if res[1].is_cuda else None
        nthr = len(grad_maxs) if (device_thr and self._thr_dev is not None) else 0
        packed = torch.cat([res[1].reshape(1), grad_maxs] +
                           ([self._thr_dev.to(torch.float32)] if nthr else []) +
                           ([flag.to(torch.float32)] if flag is not None else []))
        buf = self._host_bufs[self._nstep & 1]
        if buf is None or buf.numel() != packed.numel():
            buf = self._host_bufs[self._nstep & 1] = torch.empty(packed.numel(), dtype=torch.float32,
                                                                pin_memory=packed.is_cuda)
        self._nstep += 1
        buf.copy_(packed, non_blocking=True)
        done = None
        if packed.is_cuda:
            done = torch.cuda.Event()
            done.record()
        return res, grad_maxs, done, buf, flag is not None, nthr, device_thr

    def finish(self, pending):
        """Wait for an enqueued step; returns (calculate_loss tuple, loss, gradient maxima)."""
        res, _, done, buf, has_flag, nthr, device_thr = pending
        if done is not None:
            done.synchronize()
        host = buf.numpy().copy()
        if has_flag:
            ctc.raise_if_flagged(host[-1] != 0)
            host = host[:-1]
        if nthr:
            self.grad_max_threshs = host[-nthr:]
            host = host[:-nthr]
        if self.rolling_mads is not None and not device_thr:
            self.grad_max_threshs = self.rolling_mads.update(host[1:])
        return res, float(host[0]), host[1:]

    def __call__(self, batch_gen, sharpen=1.0, mod_factor=1.0, read_back=True):
        pending = self.enqueue(batch_gen, sharpen, mod_factor)
        if not read_back:
            return pending[0], None, pending[1]
        return self.finish(pending)
