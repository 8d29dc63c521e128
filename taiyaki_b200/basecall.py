"""Basecalling driver around the network and the inference lattice operators
 -- the per-read flow of bin/basecall.py:158-243
(`process_read`): normalise, chunk, network, posterior transition weights,
Viterbi, stitch, quality string, path -> bases.

Device-first differences from the reference, none of which change a result:
the read's signal crosses to the device once and is normalised (median / MAD by two
sorts) and chunked there; chunks of
SEVERAL reads share one batch (`process_signals`) so the recurrence kernels see
`max_concurrent_chunks` chunks even when single reads are short; stitching is
one gather per read; only the stitched path and error probabilities come back.
Beam search (`decodeutil.beamsearch`, a CPU Cython routine) is not on this path.
"""
import numpy as np
import torch

from . import basecall_helpers, qscores
from .decode import flipflop_make_trans, flipflop_viterbi
from .flipflopfings import path_to_str
from .maths import MAD_SD_FACTOR, med_mad


def med_mad_norm(x, dtype='f4'):
    """Normalise with median and MAD (bin/basecall.py:74-87)."""
    med, mad = med_mad(x)
    normed_x = (x - med) / mad
    return normed_x.astype(dtype)


def normalise_signal(signal, reverse=False, read_params=None):
    """Signal as the network sees it (bin/basecall.py:198-204)."""
    if reverse:
        signal = signal[::-1]
    if read_params is None:
        return med_mad_norm(signal)
    return ((signal - read_params['shift']) / read_params['scale']).astype('f4')


def _median_sorted(x):
    """np.median of a 1-D tensor: middle element, or (a + b) / 2 of the middle two."""
    s, _ = torch.sort(x)
    n = s.numel()
    return s[n // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2


def normalise_signal_device(signal, device, reverse=False, read_params=None):
    """`normalise_signal` on the device: the host cost of two medians per read (~1.5 ms for a
    60 k-sample read) was most of a batch's wall time.  The signal crosses as float64 --
    what `Signal.current` hands the reference -- and the arithmetic is numpy's in the same
    order (median = mean of the middle two of the sorted values; (x - med) / (factor * mad)
    in float64, then one rounding to float32), so the result is bit-identical to
    `med_mad_norm(signal.astype('f8'))`.  Returns a 1-D float32 tensor."""
    x = torch.as_tensor(np.ascontiguousarray(signal, dtype=np.float64)).to(device)
    if reverse:
        x = x.flip(0)
    if read_params is None:
        med = _median_sorted(x)
        mad = MAD_SD_FACTOR * _median_sorted((x - med).abs())
        return ((x - med) / mad).float()
    return ((x - float(read_params['shift'])) / float(read_params['scale'])).float()


def decode_chunks(trans, posterior=True, temperature=1.0):
    """Lattice decoding of chunked transition scores [T, nchunks, S]
    (bin/basecall.py:216-229 without the beam branch): returns (trans as
    decoded -- log posterior weights when `posterior` -- and the per-chunk best
    paths [T+1, nchunks])."""
    trans = trans * temperature
    if posterior:
        trans = (flipflop_make_trans(trans) + 1e-8).log()
    _, _, chunk_best_paths = flipflop_viterbi(trans)
    return trans, chunk_best_paths


def _finish_read(trans, chunk_best_paths, chunk_starts, chunk_ends, stride, alphabet,
                 fastq, qscore_scale, qscore_offset):
    """Stitch one read's chunks and turn the path into bases (+ quality string)
    (bin/basecall.py:225-243)."""
    best_path_dev = basecall_helpers.stitch_chunks(chunk_best_paths, chunk_starts, chunk_ends, stride)
    qstring = None
    if fastq:
        chunk_errprobs = qscores.errprobs_from_trans(trans, chunk_best_paths)
        errprobs = basecall_helpers.stitch_chunks(chunk_errprobs, chunk_starts, chunk_ends, stride)
        best_path = best_path_dev.cpu().numpy()
        qstring = qscores.path_errprobs_to_qstring(errprobs.cpu().numpy(), best_path,
                                                   qscore_scale, qscore_offset)
    else:
        best_path = best_path_dev.cpu().numpy()
    basecall = path_to_str(best_path, alphabet=alphabet, include_first_source=False)
    return basecall, qstring


def _run_chunks(model, chunks, n_can_state, max_concurrent_chunks):
    """Network over [chunk_size, nchunks, 1] in batches of `max_concurrent_chunks`
    (bin/basecall.py:211-214)."""
    trans = [model(some_chunks.contiguous())[:, :, :n_can_state]
             for some_chunks in torch.split(chunks, max_concurrent_chunks, 1)]
    return trans[0] if len(trans) == 1 else torch.cat(trans, 1)


def process_signal(signal, model, chunk_size, overlap, read_params, n_can_state, stride,
                   alphabet, max_concurrent_chunks, fastq=False, qscore_scale=1.0,
                   qscore_offset=0.0, beam=None, posterior=True, temperature=1.0):
    """Basecall one read's raw signal: `process_read` of bin/basecall.py:158-243
    with the signal passed in instead of read from a fast5 file.  `chunk_size`
    and `overlap` are in samples.  Returns (basecall, qstring or None, nsample)."""
    if signal is None:
        return None, None, 0
    if beam is not None:
        raise NotImplementedError('beam search decoding is not part of the device path')
    device = next(model.parameters()).device
    with torch.no_grad():
        sig = normalise_signal_device(signal, device, model.metadata['reverse'], read_params)
        chunks, chunk_starts, chunk_ends = basecall_helpers.chunk_read(sig, chunk_size, overlap)
        trans = _run_chunks(model, chunks, n_can_state, max_concurrent_chunks)
        trans, chunk_best_paths = decode_chunks(trans, posterior, temperature)
        basecall, qstring = _finish_read(trans, chunk_best_paths, chunk_starts, chunk_ends, stride,
                                         alphabet, fastq, qscore_scale, qscore_offset)
    return basecall, qstring, len(signal)


def process_signals(signals, model, chunk_size, overlap, all_read_params, n_can_state, stride,
                    alphabet, max_concurrent_chunks, fastq=False, qscore_scale=1.0,
                    qscore_offset=0.0, posterior=True, temperature=1.0):
    """Basecall several reads with their chunks sharing batches.  `signals` is a
    list of (read_id, signal); returns [(read_id, basecall, qstring, nsample)] in
    the same order, each identical to `process_signal` of that read alone (chunks
    are independent through the network and the lattice operators).  Reads shorter
    than one chunk have a different chunk length and are run by themselves."""
    device = next(model.parameters()).device
    results = [None] * len(signals)
    full = []
    with torch.no_grad():
        for i, (read_id, signal) in enumerate(signals):
            if signal is None:
                results[i] = (read_id, None, None, 0)
            elif len(signal) < chunk_size:
                results[i] = (read_id, *process_signal(
                    signal, model, chunk_size, overlap, all_read_params.get(read_id), n_can_state,
                    stride, alphabet, max_concurrent_chunks, fastq, qscore_scale, qscore_offset,
                    None, posterior, temperature))
            else:
                sig = normalise_signal_device(signal, device, model.metadata['reverse'],
                                              all_read_params.get(read_id))
                full.append((i, basecall_helpers.chunk_read(sig, chunk_size, overlap)))
        if full:
            chunks = torch.cat([c[0] for _, c in full], 1)
            trans = _run_chunks(model, chunks, n_can_state, max_concurrent_chunks)
            trans, paths = decode_chunks(trans, posterior, temperature)
            first = 0
            for i, (c, chunk_starts, chunk_ends) in full:
                n = c.shape[1]
                basecall, qstring = _finish_read(
                    trans[:, first:first + n], paths[:, first:first + n], chunk_starts, chunk_ends,
                    stride, alphabet, fastq, qscore_scale, qscore_offset)
                results[i] = (signals[i][0], basecall, qstring, len(signals[i][1]))
                first += n
    return results
