"""Training batches assembled on the device (csrc/batching.cu).

`DeviceReadStore` keeps every read's DACs, Ref_to_signal and Reference in HBM;
`sample` draws the random (read, start sample) candidates on the host exactly
as taiyaki/chunk_selection.py:57-73 / signal_mapping.py:538-548 do (a random
read, a random window inside its mapped region) and leaves everything that
touches signal or labels -- reference span, dwell filters, "first N that pass",
standardisation, stacking to [T, N, 1], flip-flop coding -- to three kernel
launches.  `prepare_random_batches` is the device twin of
training.prepare_random_batches (bin/train_flipflop.py:78-142): same tuple,
tensors already on the device.
"""
import numpy as np
import torch

from . import _lib, ctc
from .signal_mapping import Chunk

#: order of the rejection counters returned by ty_sample_chunks
REJECT_NAMES = (Chunk.rej_str_pass, Chunk.rej_str_empty_seq, Chunk.rej_str_empty_sig,
                Chunk.rej_str_short, Chunk.rej_str_null_map, Chunk.rej_str_path_buffer,
                Chunk.rej_str_mean_dwl, Chunk.rej_str_max_dwl)


class DeviceReadStore:
    def __init__(self, read_data, device):
        if len(read_data) == 0:
            raise ValueError('no reads')
        self.device = torch.device(device)
        self.nreads = len(read_data)

        def cat(arrays, dtype):
            off = np.zeros(len(arrays) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(a) for a in arrays])
            return (torch.from_numpy(np.concatenate(arrays).astype(dtype)).to(self.device),
                    torch.from_numpy(off).to(self.device))
        self.dacs, self.dacs_off = cat([r.Dacs for r in read_data], np.int16)
        self.r2s, self.r2s_off = cat([r.Ref_to_signal for r in read_data], np.int32)
        self.ref, self.ref_off = cat([r.Reference for r in read_data], np.int16)
        # current = (dacs + offset) * range / digitisation, standardised (- shift) / scale
        # (signal_mapping.py:474-477) as one multiply-add per sample
        gain = np.array([r.range / r.digitisation for r in read_data], dtype=np.float64)
        off = np.array([r.offset for r in read_data], dtype=np.float64)
        shift = np.array([r.shift_frompA for r in read_data], dtype=np.float64)
        scale = np.array([r.scale_frompA for r in read_data], dtype=np.float64)
        lin_raw = np.stack([gain, off * gain], 1)
        lin_std = np.stack([gain / scale, (off * gain - shift) / scale], 1)
        self.lin = {False: torch.from_numpy(lin_raw.astype(np.float32)).to(self.device),
                    True: torch.from_numpy(lin_std.astype(np.float32)).to(self.device)}
        regions = np.array([r.get_mapped_dacs_region() for r in read_data], dtype=np.int64)
        self.region_start, self.region_end = regions[:, 0], regions[:, 1]
        self._tables = {}

    def _label_tables(self, metadata):
        """cat-mod label -> (canonical label, mod category) tables on the device."""
        if not metadata.is_cat_mod:
            return None, None
        key = id(metadata.can_labels)
        if key not in self._tables:
            self._tables[key] = (
                torch.from_numpy(np.asarray(metadata.can_labels).astype(np.int32)).to(self.device),
                torch.from_numpy(np.asarray(metadata.mod_labels).astype(np.int32)).to(self.device))
        return self._tables[key]

    def draw_candidates(self, number_of_attempts, chunk_len, select_strands_randomly=True,
                        first_strand_index=0):
        """(read, first sample) of every attempt the reference's loop could make;
        first sample -1 marks a read too short for the chunk ('tooshort')."""
        m = int(number_of_attempts)
        if select_strands_randomly:
            reads = np.random.randint(self.nreads, size=m)
        else:
            reads = (first_strand_index + np.arange(m)) % self.nreads
        spare = self.region_end[reads] - self.region_start[reads] - chunk_len
        start = self.region_start[reads] + (np.random.random_sample(m) * np.maximum(spare, 1)).astype(np.int64)
        start = np.where(spare > 0, start, -1)
        return reads.astype(np.int32), start.astype(np.int32)

    def _staging(self, nbytes_cand, nbytes_small):
        """Pinned host buffers from a free list, returned by `finish` (allocating
        pinned memory per batch -- cudaHostAlloc -- cost milliseconds on some hosts)."""
        free = self.__dict__.setdefault('_stage_free', [])
        for i, (c, m) in enumerate(free):
            if c.numel() >= nbytes_cand and m.numel() >= nbytes_small:
                return free.pop(i)
        pin = self.device.type == 'cuda'
        return (torch.empty(max(nbytes_cand, 1 << 12), dtype=torch.uint8, pin_memory=pin),
                torch.empty(max(nbytes_small, 1 << 12), dtype=torch.uint8, pin_memory=pin))

    def launch(self, number_to_sample, chunk_len, filter_params, metadata, nbase,
               select_strands_randomly=True, first_strand_index=0, candidates=None,
               stream=None):
        """Enqueue the three batching kernels and the small device -> host copy of
        one batch on `stream` (default: the current stream) WITHOUT waiting;
        `finish(pending)` completes it.  Issuing batch k+1 before the train step of
        batch k is enqueued hides both the kernels and the read-back."""
        lib = _lib.lib()
        dev = self.device
        N, T = int(number_to_sample), int(chunk_len)
        attempts = max(N, int(N / filter_params.filter_min_pass_fraction))
        if candidates is None:
            candidates = self.draw_candidates(attempts, T, select_strands_randomly,
                                              first_strand_index)
        cand_read, cand_start = candidates
        M = len(cand_read)
        ncount = lib.ty_batch_counts_len()
        nsmall = (2 * N + 1) * 8 + ncount * 4
        cand_host, small_host = self._staging(2 * M * 4, nsmall)
        ch = cand_host[:2 * M * 4].view(torch.int32).view(2, M)
        ch[0] = torch.from_numpy(np.ascontiguousarray(cand_read, dtype=np.int32))
        ch[1] = torch.from_numpy(np.ascontiguousarray(cand_start, dtype=np.int32))
        use_filters = None not in (filter_params.median_meandwell, filter_params.mad_meandwell,
                                   filter_params.model_stride, filter_params.path_buffer)
        filt = None
        if use_filters:
            import ctypes
            filt = (ctypes.c_float * 5)(filter_params.filter_mean_dwell,
                                        filter_params.filter_max_dwell,
                                        filter_params.median_meandwell,
                                        filter_params.mad_meandwell, filter_params.path_buffer)
        can_t, mod_t = self._label_tables(metadata)
        ctx = torch.cuda.stream(stream) if stream is not None else _null_context()
        with ctx:
            cur = torch.cuda.current_stream(dev)
            cand = ch.to(dev, non_blocking=True)
            indata = torch.empty(T, N, 1, dtype=torch.float32, device=dev)
            max_seq = T          # a window of T samples spans at most T + 1 mapped positions
            seqs = torch.empty(N * (max_seq + 1), dtype=torch.int64, device=dev)
            mod_cats = torch.empty_like(seqs) if metadata.is_cat_mod else None
            # seqlen [N] | seqoff [N+1] (int64) | counters [ncount] (int32): one zero fill,
            # one device -> host copy
            small = torch.zeros(nsmall, dtype=torch.uint8, device=dev)
            small64 = small[:(2 * N + 1) * 8].view(torch.int64)
            seqlen, seqoff = small64[:N], small64[N:2 * N + 1]
            counts = small[(2 * N + 1) * 8:].view(torch.int32)
            scratch = torch.empty(3 * M + N, dtype=torch.int32, device=dev)
            rc = lib.ty_sample_chunks(
                _lib.ptr(self.dacs), _lib.ptr(self.dacs_off), _lib.ptr(self.r2s),
                _lib.ptr(self.r2s_off), _lib.ptr(self.ref), _lib.ptr(self.ref_off),
                _lib.ptr(self.lin[bool(metadata.standardize)]), _lib.ptr(cand[0]),
                _lib.ptr(cand[1]), M, N, T, filt, int(filter_params.model_stride or 0),
                int(bool(metadata.reverse)), int(nbase), _lib.ptr(can_t), _lib.ptr(mod_t),
                _lib.ptr(indata), _lib.ptr(seqs), _lib.ptr(mod_cats), _lib.ptr(seqlen),
                _lib.ptr(seqoff), _lib.ptr(counts), _lib.ptr(scratch), c_stream(cur))
            _lib.check(rc, 'ty_sample_chunks')
            _lib.count_launches(3)
            small_host[:nsmall].copy_(small, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
        return dict(N=N, ncount=ncount, indata=indata, seqs=seqs, mod_cats=mod_cats,
                    seqlen=seqlen, small_host=small_host[:nsmall], done=done,
                    stream=stream, keep=(cand, scratch, small), stage=(cand_host, small_host))

    def finish(self, pending):
        """Wait for a launched batch and hand it over on the CURRENT stream."""
        dev = self.device
        N, ncount = pending['N'], pending['ncount']
        pending['done'].synchronize()                       # the one host wait per batch
        host = pending['small_host'].numpy().copy()
        self._stage_free.append(pending['stage'])
        lens = host[:N * 8].view(np.int64)
        cnt = host[(2 * N + 1) * 8:].view(np.int32)
        n_acc, total = int(cnt[len(REJECT_NAMES)]), int(lens.sum())
        rejections = {name: int(c) for name, c in zip(REJECT_NAMES, cnt) if c}
        indata, seqs, mod_cats = pending['indata'], pending['seqs'], pending['mod_cats']
        if pending['stream'] is not None:
            # produced on the batching stream, consumed on the caller's
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(pending['done'])
            for t in (indata, seqs, mod_cats) + pending['keep']:
                if t is not None:
                    t.record_stream(cur)
        seqlens = pending['seqlen'].clone()
        if n_acc < N:       # rare: not enough chunks passed the filters
            indata = indata[:, :n_acc].contiguous()
            seqlens = seqlens[:n_acc].clone()
            lens = lens[:n_acc]
        seqs = seqs[:total]
        if mod_cats is not None:
            mod_cats = mod_cats[:total]
        ctc.hint_lengths(seqlens, int(lens.max()) if n_acc else 0, total)
        seqlens._ty_len_min = int(lens.min()) if n_acc else 0      # host copy, no sync to check
        return indata, seqs, seqlens, mod_cats, n_acc, rejections

    def sample(self, number_to_sample, chunk_len, filter_params, metadata, nbase,
               select_strands_randomly=True, first_strand_index=0, candidates=None):
        """One batch.  Returns (indata [T, N, 1] fp32, seqs int64, seqlens int64 [N],
        mod_cats int64 or None, number accepted, rejection counts) -- device
        tensors, one small device -> host copy (the counters and lengths)."""
        return self.finish(self.launch(number_to_sample, chunk_len, filter_params, metadata,
                                       nbase, select_strands_randomly, first_strand_index,
                                       candidates))


class _null_context:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def c_stream(stream):
    import ctypes
    return ctypes.c_void_p(stream.cuda_stream)


class BatchPrefetcher:
    """FIFO of batches in flight: `request` enqueues the kernels and the small
    read-back of a batch on a side stream, `next` hands the oldest one over.
    Requesting the batches of iteration k+1 before the train step of iteration k
    is enqueued puts batch assembly under the step (train_flipflop.py:78-142 is
    serial with the step in the reference).  Candidates are drawn at `request`
    time, in request order, so the numpy random stream is consumed exactly as by
    the unpipelined loop (chunk length, then candidates, per iteration)."""

    def __init__(self, store, alphabet_info, filter_params, net_info, log, use_side_stream=True):
        self.store, self.alphabet_info, self.filter_params = store, alphabet_info, filter_params
        self.net_info, self.log = net_info, log
        self.queue = []
        self.side = None
        if use_side_stream and store.device.type == 'cuda':
            self.side = torch.cuda.Stream(store.device)
            self.side.wait_stream(torch.cuda.current_stream(store.device))   # store uploads

    def request(self, chunk_len, sub_batch_size):
        self.queue.append((sub_batch_size, self.store.launch(
            sub_batch_size, chunk_len, self.filter_params, self.net_info.metadata,
            self.alphabet_info.ncan_base, stream=self.side)))

    def next(self):
        sub_batch_size, pending = self.queue.pop(0)
        out = self.store.finish(pending)
        _check_batch(out, sub_batch_size, self.log)
        return out

    def batches(self, n):
        for _ in range(n):
            yield self.next()


def _check_batch(batch, sub_batch_size, log):
    _, _, seqlens, _, n_acc, _ = batch
    if n_acc < sub_batch_size and log is not None:
        log.write(('* Warning: only {} chunks passed filters (asked for {}).\n').format(
            n_acc, sub_batch_size))
    if n_acc == 0 or seqlens._ty_len_min <= 0:
        raise Exception('Error: zero length sequence')


def prepare_random_batches(store, batch_chunk_len, sub_batch_size, target_sub_batches,
                           alphabet_info, filter_params, net_info, log,
                           select_strands_randomly=True, first_strand_index=0):
    """Device twin of training.prepare_random_batches: same tuples, tensors on the
    device, one batch at a time on the current stream (`BatchPrefetcher` is the
    pipelined form the training loop uses)."""
    total_sub_batches = 0
    while total_sub_batches < target_sub_batches:
        batch = store.sample(sub_batch_size, batch_chunk_len, filter_params, net_info.metadata,
                             alphabet_info.ncan_base, select_strands_randomly,
                             first_strand_index)
        first_strand_index += sum(batch[5].values())
        _check_batch(batch, sub_batch_size, log)
        total_sub_batches += 1
        yield batch
