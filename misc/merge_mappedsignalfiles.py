#!/usr/bin/env python3
"""Combine mapped-signal files into one -- the arguments and rules of taiyaki's
misc/merge_mappedsignalfiles.py (:13-237).  Here it is also how a multi-GPU data preparation ends:
bin/prepare_mapped_reads.py drives one GPU per process, each process remaps its share of the reads
(--input_strand_list) into its own file, and this script joins the shards.

    merge_mappedsignalfiles.py output.hdf5 --input shard0.hdf5 None --input shard1.hdf5 1000

Inputs may be per-read or batched files; a read id already copied from an earlier input is not
copied again; with a read limit for an input its reads are drawn in random order (--seed).
Alphabets must be equal, or -- with --allow_mod_merge -- compatible: the same canonical bases, and
no modified-base letter or long name standing for two different things; labels are then recoded
into the merged alphabet.  The output is the batched layout (the only one this package writes;
--batch_format is accepted for compatibility).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import alphabet  # noqa: E402
from taiyaki_b200.cmdargs import AutoBool  # noqa: E402
from taiyaki_b200.mapped_signal_files import MappedSignalReader, MappedSignalWriter  # noqa: E402

FILE_VERSION = 8        # mapped_signal_files._version


def get_parser():
    p = argparse.ArgumentParser(description='Combine mapped-signal files into a single file. '
                                'Checks that alphabets are compatible.')
    p.add_argument('output', help='Output filename')
    p.add_argument('--input', required=True, nargs=2, action='append',
                   metavar=('mapped_signal_file', 'num_reads'),
                   help='Mapped signal filename and the number of reads to merge from this file. '
                        'Specify "None" to merge all reads from a file.')
    p.add_argument('--load_in_mem', action=AutoBool, default=True,
                   help='Accepted for compatibility (input files are memory-mapped)')
    p.add_argument('--seed', default=None, type=lambda s: None if s in ('None', 'none') else int(s),
                   help='Seed for randomly selected reads when limits are set')
    p.add_argument('--allow_mod_merge', action='store_true',
                   help='Allow merging of data sets with different modified bases')
    p.add_argument('--batch_format', action='store_true',
                   help='Accepted for compatibility: the output is always the batched format')
    return p


def file_alphabets(in_fns):
    """Alphabet of every input, after checking its format version -- before any read is copied, so
    that a late mismatch does not waste a long run (merge_mappedsignalfiles.py:52-61)."""
    out = []
    for fn in in_fns:
        with MappedSignalReader(fn) as msr:
            if msr.version != FILE_VERSION:
                raise Exception('File version of mapped signal file ({}, version {}) does not match this '
                                'version of Taiyaki (file version {})'.format(fn, msr.version, FILE_VERSION))
            out.append(msr.get_alphabet_information())
    return out


def same_alphabet(a, b):
    """taiyaki/alphabet.py:240-248."""
    return (a.alphabet == b.alphabet and a.collapse_alphabet == b.collapse_alphabet and
            list(a.mod_long_names or []) == list(b.mod_long_names or []))


def assert_all_alphabets_equal(in_fns):
    infos = file_alphabets(in_fns)
    for fn, info in zip(in_fns[1:], infos[1:]):
        if not same_alphabet(infos[0], info):
            sys.stderr.write('Alphabet info in {} differs from that in {}\n'.format(fn, in_fns[0]))
            sys.exit(1)
    return infos[0]


def validate_and_merge_alphabets(in_fns):
    """Union of the inputs' modified bases over one canonical alphabet; refuses a letter that
    stands for two (canonical base, long name) pairs and a long name under two letters
    (merge_mappedsignalfiles.py:64-131)."""
    infos = file_alphabets(in_fns)
    can_bases = infos[0].can_bases
    if any(info.can_bases != can_bases for info in infos):
        sys.stderr.write('All canonical alphabets must be the same for --allow_mod_merge. Got: {}\n'.format(
            ', '.join(sorted(set(info.can_bases for info in infos)))))
        sys.exit(1)
    mods, letter_of_name, seen_in = {}, {}, {}

    def clash(fn, base, name, can, other):
        o_can, o_name = mods[other]
        sys.stderr.write('Incompatible modified bases encountered:\n\t{}={} (alt to {}) from {}\n\t'
                         '{}={} (alt to {}) from {}\n'.format(base, name, can, fn, other, o_name, o_can,
                                                              seen_in[other]))
        sys.exit(1)
    for fn, info in zip(in_fns, infos):
        for base in info.mod_bases:
            can, name = info.collapse_sequence(base), info.mod_name_conv[base]
            if base in mods:
                if mods[base] != (can, name):
                    clash(fn, base, name, can, base)
            else:
                if name in letter_of_name:
                    clash(fn, base, name, can, letter_of_name[name])
                mods[base], letter_of_name[name], seen_in[base] = (can, name), base, fn
    return alphabet.AlphabetInfo(can_bases + ''.join(mods), can_bases + ''.join(c for c, _ in mods.values()),
                                 [n for _, n in mods.values()], do_reorder=True)


def label_conversion(file_info, merged_info):
    """file label -> label of the same base in the merged alphabet."""
    return np.array([merged_info.alphabet.index(b) for b in file_info.alphabet], dtype=np.int16)


def add_file_reads(msr, msw, input_fn, allow_mod_merge, merged_info, input_limit, reads_written):
    conv = label_conversion(msr.get_alphabet_information(), merged_info) if allow_mod_merge else None
    before = len(reads_written)
    read_ids = list(msr.get_read_ids())
    if input_limit is not None:
        np.random.shuffle(read_ids)
    new_ids = [r for r in read_ids if r not in reads_written]
    if len(new_ids) < len(read_ids):
        sys.stderr.write('* {} reads found in previous file: not copying from {}.\n'.format(
            len(read_ids) - len(new_ids), input_fn))
    if input_limit is not None:
        new_ids = new_ids[:input_limit]
    for read in (msr.reads(new_ids) if new_ids else ()):
        d = read.get_read_dictionary()
        if conv is not None:
            d['Reference'] = conv[np.asarray(d['Reference'])]
        msw.write_read(d)
        reads_written.add(read.read_id)
    sys.stderr.write('Copied {} reads from {}.\n'.format(len(reads_written) - before, input_fn))
    return reads_written


def main(argv=None):
    args = get_parser().parse_args(argv)
    input_fns = [fn for fn, _ in args.input]
    input_limits = [None if n == 'None' else int(n) for _, n in args.input]
    if args.allow_mod_merge:
        merged_info = validate_and_merge_alphabets(input_fns)
        sys.stderr.write('Merged alphabet contains: {}\n'.format(str(merged_info)))
    else:
        merged_info = assert_all_alphabets_equal(input_fns)
    if args.seed is not None:
        np.random.seed(args.seed)
    reads_written = set()
    sys.stderr.write('Writing reads to {}\n'.format(args.output))
    with MappedSignalWriter(args.output, merged_info, True) as msw:
        for fn, limit in zip(input_fns, input_limits):
            with MappedSignalReader(fn) as msr:
                reads_written = add_file_reads(msr, msw, fn, args.allow_mod_merge, merged_info, limit,
                                               reads_written)
    sys.stderr.write('Copied {} reads in total.\n'.format(len(reads_written)))
    return len(reads_written)


if __name__ == '__main__':
    main()
