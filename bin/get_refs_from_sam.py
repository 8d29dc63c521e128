#!/usr/bin/env python3
"""Reference sequence of each read from its alignment -- the flow and arguments of taiyaki's
bin/get_refs_from_sam.py (:13-110): for every primary forward (flag 0) or reverse (flag 16)
alignment covering at least --min_coverage of its read, the stretch of the genomic reference
it spans (padded by --pad, reverse-complemented for the reverse strand) as a fasta record named
after the read -- the `references` input of bin/prepare_mapped_reads.py.  Host-side data
preparation: no GPU work.

The reference reads SAM or BAM through pysam, which is not in this image; here the SAM TEXT
format is parsed directly (flag, reference name, position and CIGAR are all that is used).  A BAM
file is refused with a message to convert it (`samtools view -h`).

    get_refs_from_sam.py [flags] genome.fasta alignments.sam [more.sam ...] > read_references.fasta
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200.cmdargs import AutoBool  # noqa: E402
from taiyaki_b200.prepare_mapping_funcs import fasta_file_to_dict  # noqa: E402

COMPLEMENT = str.maketrans('ATCGXNatcgxn-', 'TAGCXNtagcxn-')       # taiyaki/bio.py:12-14
CIGAR = re.compile(r'(\d+)([MIDNSHP=X])')
QUERY_ALIGNED, QUERY_CLIPPED, ON_REFERENCE = 'MI=X', 'S', 'MDN=X'


def proportion(s):
    v = float(s)
    if not 0.0 <= v <= 1.0:
        raise argparse.ArgumentTypeError('{} is not a proportion'.format(s))
    return v


def complement(seq):
    bad = set(seq) - set('ATCGXNatcgxn-')
    if bad:
        raise KeyError(sorted(bad)[0])
    return seq.translate(COMPLEMENT)


def reverse_complement(seq):
    return complement(seq)[::-1]


def get_parser():
    p = argparse.ArgumentParser(description='Extract reference sequence for each read from a SAM alignment file',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--output', default=None, metavar='filename', help='Write output to file')
    p.add_argument('--complement', default=False, action=AutoBool,
                   help='Complement all reference sequences')
    p.add_argument('--input_strand_list', default=None, help='Strand summary file containing subset')
    p.add_argument('--min_coverage', metavar='proportion', default=0.6, type=proportion,
                   help='Ignore reads with alignments shorter than min_coverage * read length')
    p.add_argument('--pad', type=int, default=0, help='Number of bases by which to pad reference sequence')
    p.add_argument('--reverse', default=False, action=AutoBool,
                   help='Reverse all reference sequences (for RNA)')
    p.add_argument('reference', help='Genomic references that reads were aligned against')
    p.add_argument('input', metavar='input.sam', nargs='+', help='SAM file(s) containing read alignments to reference')
    return p


def sam_records(sam):
    """(query name, flag, reference name, 0-based start, CIGAR operations) of each alignment line."""
    with open(sam, 'rb') as fh:
        if fh.read(2) == b'\x1f\x8b':
            raise SystemExit('{} is compressed (BAM?): this script reads SAM text; convert with '
                             '`samtools view -h`'.format(sam))
    with open(sam) as fh:
        for line in fh:
            if line.startswith('@') or not line.strip():
                continue
            f = line.rstrip('\n').split('\t')
            ops = [(int(n), op) for n, op in CIGAR.findall(f[5])] if f[5] != '*' else []
            yield f[0], int(f[1]), f[2], int(f[3]) - 1, ops


def get_refs(sam, ref_seq_dict, min_coverage=0.6, pad=0, strand_list=None):
    """(read name, its reference sequence) for the alignments that pass
    (get_refs_from_sam.py:47-80; the lengths are pysam's query_alignment_length, query_length
    and reference_end restated on the CIGAR)."""
    for name, flag, rname, start, ops in sam_records(sam):
        if flag != 0 and flag != 16:            # unmapped, secondary, supplementary, ...
            continue
        if strand_list is not None and name not in strand_list:
            continue
        aligned = sum(n for n, op in ops if op in QUERY_ALIGNED)
        query_length = aligned + sum(n for n, op in ops if op in QUERY_CLIPPED)
        if query_length == 0 or aligned / query_length < min_coverage:
            continue
        read_ref = ref_seq_dict.get(rname)
        if read_ref is None:
            continue
        end = start + sum(n for n, op in ops if op in ON_REFERENCE)
        read_ref = read_ref[max(0, start - pad):min(len(read_ref), end + pad)].upper()
        if flag == 16:
            read_ref = reverse_complement(read_ref)
        yield name, read_ref


def main(argv=None):
    args = get_parser().parse_args(argv)
    if args.output is not None and os.path.exists(args.output):
        sys.stderr.write('Output file {} already exists\n'.format(args.output))
        sys.exit(1)
    sys.stderr.write('* Loading references (this may take a while for large genomes)\n')
    references = fasta_file_to_dict(args.reference, filter_ambig=False)
    strand_list = None
    if args.input_strand_list is not None:
        with open(args.input_strand_list) as fh:
            column = fh.readline().rstrip('\n').split('\t').index('read_id')
            strand_list = frozenset(line.rstrip('\n').split('\t')[column] for line in fh if line.strip())
        sys.stderr.write('* Strand list contains {} reads\n'.format(len(strand_list)))
    sys.stderr.write('* Extracting read references using SAM alignment\n')
    fh = sys.stdout if args.output is None else open(args.output, 'w')
    count = 0
    try:
        for samfile in args.input:
            for name, read_ref in get_refs(samfile, references, args.min_coverage, args.pad, strand_list):
                if args.reverse:
                    read_ref = read_ref[::-1]
                if args.complement:
                    read_ref = complement(read_ref)
                fh.write('>{}\n{}\n'.format(name, read_ref))
                count += 1
    finally:
        if fh is not sys.stdout:
            fh.close()
    return count


if __name__ == '__main__':
    main()
