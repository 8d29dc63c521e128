#!/usr/bin/env python3
"""Basecall reads with a flip-flop model -- the flow and flags of taiyaki's
bin/basecall.py (:23-72 parser, :246-318 main) on the B200-native path
(taiyaki_b200/basecall.py: device chunking, network, posterior transition
weights, Viterbi, stitching).

    basecall.py [flags] input_folder model.checkpoint > calls.fa

`input_folder` is a directory of fast5 files (single- or multi-read, or one such file) as in
the reference -- decoded by taiyaki_b200/fast5utils.py, this image has no ont_fast5_api -- whose
reads are called from their current in pA (Signal(read).current, bin/basecall.py:92-116); or it
holds one `<read_id>.npy` per read (1-D array of current in pA, or raw DACs if --scaling
gives shift / scale for the read), or is a single `.npz` whose keys are read ids.  One process drives the GPU; instead of a pool of
worker processes (--jobs), chunks of several reads share each batch
(--reads_per_batch).  Beam search (--beam) is not supported.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import basecall, basecall_helpers, fast5utils, helpers  # noqa: E402
from taiyaki_b200.cmdargs import (AutoBool, DeviceAction, FileExists, Maybe, NonNegative,  # noqa: E402
                                  Positive)
from taiyaki_b200.flipflopfings import nstate_flipflop  # noqa: E402
from taiyaki_b200.prepare_mapping_funcs import get_per_read_params_dict_from_tsv  # noqa: E402
from taiyaki_b200.signal import Signal  # noqa: E402


def get_parser():
    p = argparse.ArgumentParser(description='Basecall reads using a taiyaki model',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--alphabet', default='ACGT')
    p.add_argument('--device', default='cuda:0', action=DeviceAction,
                   help='GPU to use: an integer, "cuda:2", "cuda2" or "cuda" (this path has no CPU mode)')
    p.add_argument('--limit', default=None, type=Maybe(Positive(int)), help='Limit number of reads to process')
    p.add_argument('--output', default=None, help='Write output to file (default stdout)')
    p.add_argument('--quiet', default=False, action=AutoBool, help="Don't print progress information to stdout")
    p.add_argument('--input_strand_list', default=None, action=FileExists,
                   help='File with a read_id column: only these reads are called')
    p.add_argument('--chunk_size', type=Positive(int), metavar='blocks',
                   default=basecall_helpers._DEFAULT_CHUNK_SIZE,
                   help='Size of signal chunks sent to GPU is chunk_size * model stride')
    p.add_argument('--fastq', default=False, action=AutoBool,
                   help='Write output in fastq format (default is fasta)')
    p.add_argument('--max_concurrent_chunks', type=Positive(int), default=128,
                   help='Maximum number of chunks to call at once')
    p.add_argument('--reads_per_batch', type=Positive(int), default=16,
                   help='Reads whose chunks are pooled into shared batches')
    p.add_argument('--overlap', type=NonNegative(int), metavar='blocks',
                   default=basecall_helpers._DEFAULT_OVERLAP,
                   help='Overlap between signal chunks sent to GPU')
    p.add_argument('--posterior', default=True, action=AutoBool,
                   help='Use posterior-viterbi decoding')
    p.add_argument('--qscore_offset', type=float, default=0.0)
    p.add_argument('--qscore_scale', type=float, default=1.0)
    p.add_argument('--reverse', default=False, action=AutoBool,
                   help='Reverse sequences in output')
    p.add_argument('--scaling', default=None, action=FileExists, help='Path to TSV containing per-read scaling params')
    p.add_argument('--temperature', default=1.0, type=float,
                   help='Scaling factor applied to network outputs before decoding')
    p.add_argument('--recursive', default=True, action=AutoBool,
                   help='Search for fast5s recursively within input_folder')
    p.add_argument('input_folder', help='Directory containing single or multi-read fast5 files '
                                        '(or <read_id>.npy signals, or one .npz)')
    p.add_argument('model', help='Model checkpoint file to use for basecalling')
    return p


def get_signal(read_filename, read_id, loader=None):
    """Current in pA of one read of a fast5 file, None when it cannot be read
    (bin/basecall.py:92-116).  `loader`: a fast5utils.ReadLoader shared by consecutive reads."""
    try:
        if loader is not None:
            return Signal(loader.get_read(read_filename, read_id)).current
        with fast5utils.get_fast5_file(read_filename, 'r') as f5file:
            return Signal(f5file.get_read(read_id)).current
    except Exception as e:
        sys.stderr.write('Unable to obtain signal for {} from {}.\n{}\n'.format(
            read_id, read_filename, repr(e)))
        return None


def _is_array_input(input_folder):
    if os.path.isfile(input_folder):
        return input_folder.endswith('.npz')
    return any(fn.endswith('.npy') for fn in os.listdir(input_folder))


def iterate_signals(input_folder, limit=None, strand_list=None, recursive=True):
    """Yield (read_id, signal) from fast5 files, a folder of .npy files or one .npz."""
    if not _is_array_input(input_folder):
        with fast5utils.ReadLoader() as loader:
            for filename, read_id in fast5utils.iterate_fast5_reads(
                    input_folder, limit=limit, strand_list=strand_list, recursive=recursive):
                yield read_id, get_signal(filename, read_id, loader)
        return
    keep = None
    if strand_list is not None:
        with open(strand_list) as fh:
            header = fh.readline().rstrip('\n').split('\t')
            c = header.index('read_id')
            keep = frozenset(line.rstrip('\n').split('\t')[c] for line in fh)
    n = 0
    if os.path.isfile(input_folder):
        with np.load(input_folder) as z:
            for read_id in z.files:
                if keep is not None and read_id not in keep:
                    continue
                if limit is not None and n >= limit:
                    return
                n += 1
                yield read_id, z[read_id]
        return
    for fn in sorted(os.listdir(input_folder)):
        if not fn.endswith('.npy'):
            continue
        read_id = fn[:-4]
        if keep is not None and read_id not in keep:
            continue
        if limit is not None and n >= limit:
            return
        n += 1
        try:
            yield read_id, np.load(os.path.join(input_folder, fn))
        except Exception as e:
            sys.stderr.write('Unable to obtain signal for {} from {}.\n{}\n'.format(
                read_id, fn, repr(e)))
            yield read_id, None


def main(argv=None):
    args = get_parser().parse_args(argv)
    import torch
    all_read_params = {}
    if args.scaling is not None:
        sys.stderr.write('* Loading read scaling parameters from {}.\n'.format(args.scaling))
        all_read_params = get_per_read_params_dict_from_tsv(args.scaling)
    device = torch.device(args.device)
    torch.cuda.set_device(device)
    model = helpers.load_model(args.model).to(device)
    stride = helpers.guess_model_stride(model)
    chunk_size = args.chunk_size * stride
    overlap = args.overlap * stride
    n_can_state = nstate_flipflop(len(args.alphabet))

    sys.stderr.write('* Calling reads.\n')
    nbase, ncalled, nread, nsample = 0, 0, 0, 0
    t0 = time.time()
    startcharacter = '@' if args.fastq else '>'
    fh = sys.stdout if args.output is None else open(args.output, 'w')

    def flush(pending):
        nonlocal nbase, ncalled, nread, nsample
        for read_id, call, qstring, read_nsample in basecall.process_signals(
                pending, model, chunk_size, overlap, all_read_params, n_can_state, stride,
                args.alphabet, args.max_concurrent_chunks, args.fastq, args.qscore_scale,
                args.qscore_offset, args.posterior, args.temperature):
            if call is not None and len(call) > 0:
                fh.write('{}{}\n{}\n'.format(startcharacter, read_id,
                                             call[::-1] if args.reverse else call))
                nbase += len(call)
                ncalled += 1
                if args.fastq:
                    fh.write('+\n{}\n'.format(qstring[::-1] if args.reverse else qstring))
            nread += 1
            nsample += read_nsample

    pending = []
    for rec in iterate_signals(args.input_folder, args.limit, args.input_strand_list, args.recursive):
        if args.scaling is not None and rec[0] not in all_read_params:
            continue
        pending.append(rec)
        if len(pending) >= args.reads_per_batch:
            flush(pending)
            pending = []
    if pending:
        flush(pending)
    if fh is not sys.stdout:
        fh.close()
    total_time = time.time() - t0
    sys.stderr.write('* Called {} reads in {:.2f}s\n'.format(nread, int(total_time)))
    sys.stderr.write('* {:7.2f} kbase / s\n'.format(nbase / total_time / 1000.0))
    sys.stderr.write('* {:7.2f} ksample / s\n'.format(nsample / total_time / 1000.0))
    sys.stderr.write('* {} reads failed.\n\n'.format(nread - ncalled))


if __name__ == '__main__':
    main()
