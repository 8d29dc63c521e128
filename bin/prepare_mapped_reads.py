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
One process drives one GPU; --jobs is replaced by --reads_per_batch (reads aligned per
launch).  The output is the batched mapped-signal format (read back by
bin/train_flipflop.py).

Several GPUs: reads are independent, so the job shards by read with no exchange between
devices.  `torchrun --nproc-per-node G bin/prepare_mapped_reads.py ...` gives rank r the reads
whose position in the input order is r modulo G on device cuda:LOCAL_RANK; each rank writes
`<output>.shard<r>of<G>`, and after a barrier rank 0 joins the shards into `<output>` and
removes them.  `--shard r G` does one share by hand (another node, another time) and leaves its
file for misc/merge_mappedsignalfiles.py.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import alphabet, fast5utils, helpers  # noqa: E402
from taiyaki_b200.cmdargs import AutoBool, DeviceAction, FileExists, Maybe, Positive  # noqa: E402
from taiyaki_b200.signal import Signal  # noqa: E402
from taiyaki_b200.prepare_mapping_funcs import (  # noqa: E402
    fasta_file_to_dict, generate_output_from_results, get_per_read_params_dict_from_tsv,
    remap_reads)


def get_parser():
    p = argparse.ArgumentParser(
        description='Prepare data for model training and save to hdf5 file by remapping with '
        'flip-flop model', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--alphabet', default='ACGT')
    p.add_argument('--device', default='cuda:0', action=DeviceAction,
                   help='GPU to use: an integer, "cuda:2", "cuda2" or "cuda" (this path has no CPU mode)')
    p.add_argument('--input_strand_list', default=None, action=FileExists)
    p.add_argument('--limit', default=None, type=Maybe(Positive(int)))
    p.add_argument('--overwrite', default=False, action=AutoBool, help='Whether to overwrite any output files')
    p.add_argument('--reads_per_batch', default=64, type=Positive(int), help='Reads aligned per launch')
    p.add_argument('--localpen', metavar='penalty', default=0.0, type=float,
                   help='Penalty for local mapping')
    p.add_argument('--max_read_length', metavar='bases', default=None, type=Maybe(int),
                   help="Don't attempt remapping for reads longer than this")
    p.add_argument('--mod', nargs=3, metavar=('mod_base', 'canonical_base', 'mod_long_name'),
                   default=[], action='append', help='Modified base description')
    p.add_argument('--shard', nargs=2, type=int, default=None, metavar=('index', 'count'),
                   help='Remap only the reads whose position in the input order is index modulo count, '
                        'into <output>.shard<index>of<count> (set from RANK / WORLD_SIZE under torchrun)')
    p.add_argument('--recursive', default=True, action=AutoBool,
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


def in_shard(index, shard):
    return shard is None or index % shard[1] == shard[0]


def iterate_npz_reads(input_folder, limit=None, strand_list=None, shard=None):
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
        if not in_shard(n - 1, shard):
            continue
        with np.load(os.path.join(input_folder, fn)) as z:
            yield {'read_id': read_id, 'dacs': z['dacs'], 'offset': float(z['offset']),
                   'range': float(z['range']), 'digitisation': float(z['digitisation'])}


def iterate_fast5_reads(input_folder, limit=None, strand_list=None, recursive=True, wanted=None,
                        shard=None):
    """Raw reads of the fast5 files below `input_folder` (fast5utils.iterate_fast5_reads: strand
    list, limit) as the dictionaries remap_reads takes.  `wanted(read_id)` False skips loading the
    samples of a read that will be rejected anyway; a read whose samples cannot be loaded is
    passed on with dacs None and reported as READ_ID_INFO_NOT_FOUND
    (prepare_mapping_funcs.py:62-68)."""
    with fast5utils.ReadLoader() as loader:
        for index, (filename, read_id) in enumerate(fast5utils.iterate_fast5_reads(
                input_folder, limit=limit, strand_list=strand_list, recursive=recursive)):
            if not in_shard(index, shard):
                continue
            read = {'read_id': read_id, 'dacs': None, 'offset': 0.0, 'range': 1.0, 'digitisation': 1.0}
            if wanted is None or wanted(read_id):
                try:
                    sig = Signal(loader.get_read(filename, read_id))
                    read.update(dacs=sig.untrimmed_dacs, offset=float(sig.offset), range=float(sig.range),
                                digitisation=float(sig.digitisation))
                except Exception as e:
                    sys.stderr.write('Unable to obtain signal for {} from {}.\n{}\n'.format(
                        read_id, filename, repr(e)))
            yield read


def iterate_raw_reads(input_folder, limit=None, strand_list=None, recursive=True, wanted=None,
                      shard=None):
    """`shard` = (index, count): only the reads at positions index modulo count of the (limited,
    strand-list filtered) input order; the samples of the others are not loaded."""
    if os.path.isdir(input_folder) and any(fn.endswith('.npz') for fn in os.listdir(input_folder)):
        return iterate_npz_reads(input_folder, limit, strand_list, shard)
    return iterate_fast5_reads(input_folder, limit, strand_list, recursive, wanted, shard)


def shard_of_process(args, environ=os.environ):
    """((index, count) or None, launched by torchrun?)."""
    if args.shard is not None:
        index, count = args.shard
        if not 0 <= index < count:
            raise SystemExit('--shard index count: need 0 <= index < count')
        return ((index, count) if count > 1 else None), False
    world = int(environ.get('WORLD_SIZE', '1'))
    if world > 1:
        return (int(environ['RANK']), world), True
    return None, False


def shard_filename(output, shard):
    return '{}.shard{}of{}'.format(output, *shard)


def join_shards(output, shard_files, alphabet_info):
    """All reads of the shard files, in shard order, into `output`; the shard files are removed."""
    from taiyaki_b200.mapped_signal_files import MappedSignalReader, MappedSignalWriter
    count = 0
    with MappedSignalWriter(output, alphabet_info) as msw:
        for fn in shard_files:
            with MappedSignalReader(fn) as msr:
                for read in msr.reads():
                    msw.write_read(read.get_read_dictionary())
                    count += 1
    for fn in shard_files:
        os.remove(fn)
    return count


def main(argv=None):
    args = get_parser().parse_args(argv)
    print('Running prepare_mapping using flip-flop remapping')
    shard, launched = shard_of_process(args)
    output = args.output if shard is None else shard_filename(args.output, shard)
    # under torchrun rank 0 also writes the joined file; a share done by hand (--shard) only its own
    for fn in ({args.output, output} if launched else {output}):
        if not args.overwrite and os.path.exists(fn):
            print('Cowardly refusing to overwrite {}'.format(fn))
            sys.exit(1)
    alphabet_info = make_alphabet_info(args.alphabet, args.mod)
    print('Converting references to labels using {}'.format(str(alphabet_info)))
    import torch
    device = args.device
    if launched and 'LOCAL_RANK' in os.environ and device == get_parser().get_default('device'):
        device = 'cuda:{}'.format(os.environ['LOCAL_RANK'])         # one process per GPU
    device = torch.device(device)
    torch.cuda.set_device(device)
    if launched:
        import torch.distributed as dist
        dist.init_process_group('gloo')          # barriers only: no data crosses between the shards
    per_read_params_dict = get_per_read_params_dict_from_tsv(args.input_per_read_params)
    model = helpers.load_model(args.model).to(device)
    stride = helpers.guess_model_stride(model)
    # references with a letter outside the alphabet are dropped, their reads reported as
    # NO_REF_FOUND (bin/prepare_mapped_reads.py:119-120)
    references = fasta_file_to_dict(args.references, alphabet=alphabet_info.alphabet)

    def results():
        def wanted(read_id):      # signals of reads without reference or parameters are not loaded
            return read_id in references and read_id in per_read_params_dict
        pending = []
        for read in iterate_raw_reads(args.input_folder, args.limit, args.input_strand_list,
                                      args.recursive, wanted, shard):
            read['ref'] = references.get(read['read_id'])
            pending.append(read)
            if len(pending) >= args.reads_per_batch:
                yield from remap_reads(pending, model, per_read_params_dict, alphabet_info,
                                       args.max_read_length, args.localpen, stride)
                pending = []
        if pending:
            yield from remap_reads(pending, model, per_read_params_dict, alphabet_info,
                                   args.max_read_length, args.localpen, stride)

    done = generate_output_from_results(results(), output, alphabet_info)
    if launched:
        dist.barrier()
        if shard[0] == 0:
            total = join_shards(args.output, [shard_filename(args.output, (r, shard[1]))
                                              for r in range(shard[1])], alphabet_info)
            sys.stderr.write('* {} reads from {} shards joined into {}\n'.format(total, shard[1], args.output))
        dist.barrier()
        dist.destroy_process_group()
    return done


if __name__ == '__main__':
    main()
