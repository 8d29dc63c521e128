"""Chunking and stitching around the network for basecalling -- the interface of
taiyaki/basecall_helpers.py (chunk_read :11-45, stitch_chunks :48-99, run_model
:102-171), laid out for the device: the chunk tensor is ONE
gather of the read's signal and the stitched output ONE gather of the chunked
output, instead of a Python loop of slice copies per chunk.  Both functions
accept host arrays / tensors as well (same arithmetic, used by the CPU tests)."""
import numpy as np
import torch

from .helpers import get_model_device, guess_model_stride

_DEFAULT_CHUNK_SIZE = 1000
_DEFAULT_OVERLAP = 100


def chunk_bounds(nsample, chunk_size, overlap):
    """Start / end (exclusive) sample of every chunk (basecall_helpers.py:30-36):
    chunks end every `chunk_size - overlap` samples, the last one ends with the
    read; a read shorter than `chunk_size` is a single short chunk."""
    if nsample < chunk_size:
        return np.array([0]), np.array([nsample])
    chunk_ends = np.arange(chunk_size, nsample, chunk_size - overlap, dtype=int)
    chunk_ends = np.concatenate([chunk_ends, [nsample]], 0)
    return chunk_ends - chunk_size, chunk_ends


def chunk_read(signal, chunk_size, overlap):
    """Divide `signal` into overlapping chunks (basecall_helpers.py:11-45).

    Returns (chunks [chunk_size, nchunks, 1] float32, chunk_starts, chunk_ends).
    A numpy signal gives a numpy chunk array as in the reference; a torch tensor
    (host or device) gives a tensor on the same device, built with one gather."""
    chunk_starts, chunk_ends = chunk_bounds(len(signal), chunk_size, overlap)
    if len(signal) < chunk_size:
        return signal[:, None, None], chunk_starts, chunk_ends
    if isinstance(signal, torch.Tensor):
        idx = (torch.arange(chunk_size, device=signal.device)[:, None]
               + torch.as_tensor(chunk_starts, device=signal.device)[None, :])
        return signal.float()[idx][:, :, None], chunk_starts, chunk_ends
    idx = np.arange(chunk_size)[:, None] + chunk_starts[None, :]
    return np.asarray(signal)[idx][:, :, None].astype('f4'), chunk_starts, chunk_ends


def stitch_ranges(chunk_starts, chunk_ends, stride, path_stitching=False):
    """Block range [start, end) kept from every chunk's output
    (basecall_helpers.py:68-97): neighbouring chunks meet in the middle of their
    overlap; `path_stitching` shifts the ranges of a (T+1)-long path by one."""
    cs = np.asarray(chunk_starts, dtype=np.int64)
    ce = np.asarray(chunk_ends, dtype=np.int64)
    n = len(cs)
    start = np.empty(n, dtype=np.int64)
    end = np.empty(n, dtype=np.int64)
    start[0] = cs[0] // stride
    start[1:] = (ce[:-1] - cs[1:]) // (2 * stride)
    end[:-1] = (ce[:-1] + cs[1:] - 2 * cs[:-1]) // (2 * stride)
    end[-1] = (ce[-1] - cs[-1]) // stride
    if path_stitching:
        start[1:] += 1
        end += 1
    return start, end


def stitch_chunks(out, chunk_starts, chunk_ends, stride, path_stitching=False):
    """Stitch network output or Viterbi paths of overlapping chunks
    (basecall_helpers.py:48-99): `out` is [time, chunks, ...]; returns the
    [blocks, ...] tensor of the whole read."""
    nchunks = out.shape[1]
    if nchunks == 1:
        return out[:, 0]
    start, end = stitch_ranges(chunk_starts, chunk_ends, stride, path_stitching)
    end = np.minimum(end, out.shape[0])          # python slices clip at the end
    count = np.maximum(end - start, 0)
    chunk_idx = np.repeat(np.arange(nchunks), count)
    first = np.cumsum(count) - count
    time_idx = np.arange(int(count.sum())) - np.repeat(first - start, count)
    t = torch.as_tensor(time_idx, device=out.device)
    c = torch.as_tensor(chunk_idx, device=out.device)
    return out[t, c]


def run_model(normed_signal, model, chunk_size=_DEFAULT_CHUNK_SIZE, overlap=_DEFAULT_OVERLAP,
              max_concur_chunks=None, return_numpy=True, return_tensor_on_device=True):
    """Hook for megalodon (basecall_helpers.py:102-171): chunk, run the network,
    stitch.  `chunk_size` and `overlap` are in blocks (multiples of the stride).
    The signal crosses to the device once; chunking and stitching run there."""
    device = get_model_device(model)
    stride = guess_model_stride(model)
    chunk_size *= stride
    overlap *= stride
    signal = torch.as_tensor(np.ascontiguousarray(normed_signal), dtype=torch.float32).to(device)
    chunks, chunk_starts, chunk_ends = chunk_read(signal, chunk_size, overlap)
    with torch.no_grad():
        if max_concur_chunks is None:
            out = model(chunks)
        else:
            out = torch.cat([model(some_chunks.contiguous())
                             for some_chunks in torch.split(chunks, max_concur_chunks, 1)], 1)
        stitched_chunks = stitch_chunks(out, chunk_starts, chunk_ends, stride)
    if return_numpy:
        return stitched_chunks.cpu().numpy()
    if return_tensor_on_device:
        return stitched_chunks
    return stitched_chunks.cpu()
