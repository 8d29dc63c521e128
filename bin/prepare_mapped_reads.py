#!/usr/bin/env python3
"""Prepare training data by remapping reads to their references with a flip-flop model --
the flow and arguments of taiyaki's bin/prepare_mapped_reads.py (:17-140) on the B200-native
path (taiyaki_b200/prepare_mapping_funcs.py: network over the whole read, ONE alignment
launch per group of reads, csrc/remap.cu).

    prepare_mapped_reads.py [flags] input_folder per_read_params.tsv output.hdf5 \\
        model.checkpoint references.fasta

`input_folder` is a directory of fast5 files, single- or multi-read, as in the reference
(decoded by taiyaki_b200/fast5utils.py over the package's plain-Python HDF5 reader -- this image
has no ont_fast5_api / h5py), or a directory with one `<read_id>.npz` per read holding `dacs`
(raw int16 samples), `offset`, `range`, `digitisation`.
One process drives the GPU; --jobs is replaced by --reads_per_batch (reads aligned per
launch).  The output is the batched mapped-signal format (read back by
bin/train_flipflop.py).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import alphabet, fast5utils, helpers  # noqa: E402
from taiyaki_b200.signal import Signal  # noqa: E402
from taiyaki_b200.prepare_mapping_funcs import (  # noqa: E402
    fasta_file_to_dict, generate_output_from_results, get_per_read_params_dict_from_tsv,
    remap_reads)


def get_parser():
    p = argparse.ArgumentParser(
        description='Prepare data for model training and save to hdf5 file by remapping with '
        'flip-flop model', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--alphabet', default='ACGT')
    p.add_argument('--device', default='cuda:0')
    p.add_argument('--input_strand_list', default=None)
    p.add_argument('--limit', default=None, type=int)
    p.add_argument('--overwrite', default=False, action='store_true')
    p.add_argument('--reads_per_batch', default=64, type=int, help='Reads aligned per launch')
    p.add_argument('--localpen', metavar='penalty', default=0.0, type=float,
                   help='Penalty for local mapping')
    p.add_argument('--max_read_length', metavar='bases', default=None,
                   type=lambda s: None if s in ('None', 'none') else int(s),
                   help="Don't attempt remapping for reads longer than this")
    p.add_argument('--mod', nargs=3, metavar=('mod_base', 'canonical_base', 'mod_long_name'),
                   default=[], action='append', help='Modified base description')
    p.add_argument('--recursive', default=True, nargs='?', const=True,
                   type=lambda v: str(v).lower() in ('1', 'true', 'yes', 'on'),
                   help='Search for fast5s recursively within input_folder')
    p.add_argument('input_folder', help='Directory containing single or multi-read fast5 files '
                                        '(or <read_id>.npz raw reads)')
    p.add_argument('input_per_read_params', help='Input per read parameter .tsv file')
    p.add_argument('output', help='Output HDF5 file')
    p.add_argument('model', help='Taiyaki model file')
    p.add_argument('references', help='Single fasta file containing references for each read')
    return p


def make_alphabet_info(canonical, mods):
    """bin/prepare_mapped_reads.py:75-95."""
    modified_bases = [elt[0] for elt in mods]
    canonical_bases = [elt[1] for elt in mods]
    for b in modified_bases:
        assert len(b) == 1, 'Modified bases must be a single character, got {}'.format(b)
        assert b not in canonical, 'Modified base must not be a canonical base, got {}'.format(b)
    for b in canonical_bases:
        assert len(b) == 1, ('Canonical coding for modified bases must be a single character, '
                             'got {}').format(b)
        assert b in canonical, ('Canonical coding for modified base must be a canonical base, '
                                'got {})').format(b)
    return alphabet.AlphabetInfo(canonical + ''.join(modified_bases),
                                 canonical + ''.join(canonical_bases),
                                 [elt[2] for elt in mods], do_reorder=True)


def iterate_npz_reads(input_folder, limit=None, strand_list=None):
    keep = None
    if strand_list is not None:
        with open(strand_list) as fh:
            header = fh.readline().rstrip('\n').split('\t')
            c = header.index('read_id')
            keep = frozenset(line.rstrip('\n').split('\t')[c] for line in fh)
    n = 0
    for fn in sorted(os.listdir(input_folder)):
        if not fn.endswith('.npz'):
            continue
        read_id = fn[:-4]
        if keep is not None and read_id not in keep:
            continue
        if limit is not None and n >= limit:
            return
        n += 1
        with np.load(os.path.join(input_folder, fn)) as z:
            yield {'read_id': read_id, 'dacs': z['dacs'], 'offset': float(z['offset']),
                   'range': float(z['range']), 'digitisation': float(z['digitisation'])}


def iterate_fast5_reads(input_folder, limit=None, strand_list=None, recursive=True, wanted=None):
    """Raw reads of the fast5 files below `input_folder` (fast5utils.iterate_fast5_reads: strand
    list, limit) as the dictionaries remap_reads takes.  `wanted(read_id)` False skips loading the
    samples of a read that will be rejected anyway; a read whose samples cannot be loaded is
    passed on with dacs None and reported as READ_ID_INFO_NOT_FOUND
    (prepare_mapping_funcs.py:62-68)."""
    for filename, read_id in fast5utils.iterate_fast5_reads(
            input_folder, limit=limit, strand_list=strand_list, recursive=recursive):
        read = {'read_id': read_id, 'dacs': None, 'offset': 0.0, 'range': 1.0, 'digitisation': 1.0}
        if wanted is None or wanted(read_id):
            try:
                with fast5utils.get_fast5_file(filename, 'r') as f5file:
                    sig = Signal(f5file.get_read(read_id))
                read.update(dacs=sig.untrimmed_dacs, offset=float(sig.offset), range=float(sig.range),
                            digitisation=float(sig.digitisation))
            except Exception as e:
                sys.stderr.write('Unable to obtain signal for {} from {}.\n{}\n'.format(
                    read_id, filename, repr(e)))
        yield read


def iterate_raw_reads(input_folder, limit=None, strand_list=None, recursive=True, wanted=None):
    if os.path.isdir(input_folder) and any(fn.endswith('.npz') for fn in os.listdir(input_folder)):
        return iterate_npz_reads(input_folder, limit, strand_list)
    return iterate_fast5_reads(input_folder, limit, strand_list, recursive, wanted)


def main(argv=None):
    args = get_parser().parse_args(argv)
    print('Running prepare_mapping using flip-flop remapping')
    if not args.overwrite and os.path.exists(args.output):
        print('Cowardly refusing to overwrite {}'.format(args.output))
        sys.exit(1)
    alphabet_info = make_alphabet_info(args.alphabet, args.mod)
    print('Converting references to labels using {}'.format(str(alphabet_info)))
    import torch
    device = torch.device(args.device)
    torch.cuda.set_device(device)
    per_read_params_dict = get_per_read_params_dict_from_tsv(args.input_per_read_params)
    model = helpers.load_model(args.model).to(device)
    stride = helpers.guess_model_stride(model)
    references = fasta_file_to_dict(args.references)

    def results():
        pending = []
        def wanted(read_id):      # signals of reads without reference or parameters are not loaded
            return read_id in references and read_id in per_read_params_dict
        for read in iterate_raw_reads(args.input_folder, args.limit, args.input_strand_list,
                                      args.recursive, wanted):
            read['ref'] = references.get(read['read_id'])
            pending.append(read)
            if len(pending) >= args.reads_per_batch:
                yield from remap_reads(pending, model, per_read_params_dict, alphabet_info,
                                       args.max_read_length, args.localpen, stride)
                pending = []
        if pending:
            yield from remap_reads(pending, model, per_read_params_dict, alphabet_info,
                                   args.max_read_length, args.localpen, stride)

    return generate_output_from_results(results(), args.output, alphabet_info)


if __name__ == '__main__':
    main()
