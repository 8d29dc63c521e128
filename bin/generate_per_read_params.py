#!/usr/bin/env python3
"""Per-read scaling parameters from raw reads -- the flow and arguments of taiyaki's
bin/generate_per_read_params.py (:17-100): for every read of a directory of fast5 files
(single- or multi-read, optionally selected by a strand list) one row

    UUID  trim_start  trim_end  shift  scale

with shift / scale = median / MAD-derived spread of the read's current in pA
(maths.med_mad of Signal(read).current: the whole read -- the trim values are passed on to
the consumers of the table, they do not enter the estimate) -- the table
bin/prepare_mapped_reads.py and bin/basecall.py --scaling take.  Host-side data preparation:
no GPU work in this script.

    generate_per_read_params.py [flags] input_folder > read_params.tsv
"""
import argparse
import csv
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import fast5utils  # noqa: E402
from taiyaki_b200.cmdargs import AutoBool  # noqa: E402
from taiyaki_b200.maths import med_mad  # noqa: E402
from taiyaki_b200.signal import Signal  # noqa: E402


def non_negative_int(s):
    v = int(s)
    if v < 0:
        raise argparse.ArgumentTypeError('{} is negative'.format(s))
    return v


def get_parser():
    p = argparse.ArgumentParser(description='Shift and scale of every read from the median / MAD of its current',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--input_strand_list', default=None,
                   help='Strand list TSV file with columns filename_fast5 or read_id or both')
    p.add_argument('--limit', default=None, type=lambda s: None if s in ('None', 'none') else int(s),
                   help='Limit number of reads to process')
    p.add_argument('--output', default=None, metavar='filename', help='Write output to file')
    p.add_argument('--recursive', default=True, action=AutoBool,
                   help='Search for fast5s recursively within input_folder')
    p.add_argument('--jobs', default=1, type=int, help='Accepted for compatibility; reads are processed in turn')
    p.add_argument('--trim', default=(200, 50), nargs=2, type=non_negative_int, metavar=('beginning', 'end'),
                   help='Number of samples to trim off start and end')
    p.add_argument('input_folder', help='Directory containing single or multi-read fast5 files')
    return p


def one_read_shift_scale(read_tuple, loader=None):
    """(read id, shift, scale) of one (file, read id) pair; (None, None, None) when the signal
    cannot be read, NaNs for an empty signal (generate_per_read_params.py:33-76)."""
    read_filename, read_id = read_tuple
    try:
        if loader is not None:
            sig = Signal(loader.get_read(read_filename, read_id))
        else:
            with fast5utils.get_fast5_file(read_filename, 'r') as f5file:
                sig = Signal(f5file.get_read(read_id))
    except Exception as e:
        sys.stderr.write('Unable to obtain signal for {} from {}.\n{}\n'.format(read_id, read_filename, repr(e)))
        return None, None, None
    current = sig.current
    if len(current) == 0:
        return read_id, np.nan, np.nan
    shift, scale = med_mad(current)
    return read_id, shift, scale


def main(argv=None):
    args = get_parser().parse_args(argv)
    if args.output is not None and os.path.exists(args.output):
        sys.stderr.write('Output file {} already exists\n'.format(args.output))
        sys.exit(1)
    trim_start, trim_end = args.trim
    reads = fast5utils.iterate_fast5_reads(args.input_folder, limit=args.limit,
                                           strand_list=args.input_strand_list, recursive=args.recursive)
    fh = sys.stdout if args.output is None else open(args.output, 'w')
    nrow = 0
    loader = fast5utils.ReadLoader()
    try:
        writer = csv.writer(fh, delimiter='\t', lineterminator='\n')
        writer.writerow(['UUID', 'trim_start', 'trim_end', 'shift', 'scale'])
        for result in (one_read_shift_scale(r, loader) for r in reads):
            if all(result):         # as the reference: drops unreadable reads (and a shift of exactly 0)
                read_id, shift, scale = result
                writer.writerow([read_id, trim_start, trim_end, shift, scale])
                nrow += 1
    finally:
        loader.close()
        if fh is not sys.stdout:
            fh.close()
    return nrow


if __name__ == '__main__':
    main()
