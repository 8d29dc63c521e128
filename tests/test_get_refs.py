"""bin/get_refs_from_sam.py: per-read reference sequences from SAM alignments (the `references`
input of bin/prepare_mapped_reads.py), restating bin/get_refs_from_sam.py:47-80 of the reference on
the SAM text format -- pysam's query_alignment_length / query_length / reference_start /
reference_end are CIGAR arithmetic, written out in the test.

Parity note: pysam is not in the build container, so the reference script cannot be run for golden
output, and the reference's fixture test/data/per_read_references.fasta was made by an earlier
version of the script (its records extend over the soft-clipped ends of the reads: 89 + 1 bases
more than the aligned span for read db6b45aa) -- it is not what the current script prints.  The
tests therefore pin the rules on hand-made alignments and check the reference's SAM fixtures
against the genome directly."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = '/root/reference/test/data'
needs_ref_files = pytest.mark.skipif(not os.path.isdir(REF_DATA),
                                     reason='the reference tree is only in the build container')
GENOME = 'ACGTTGCAAGGCTTAACCGGATCGATCGTTAGCAGGCTAGCTAACGGATTCAGGCTA'          # 57 bases
SAM = '\n'.join([
    '@HD\tVN:1.6', '@SQ\tSN:chr\tLN:57',
    # forward, 2 soft-clipped + 10 aligned (8M 1I 1M ... ): reference span 8 + 2 (D) + 1 = 11 from position 5
    'fwd\t0\tchr\t5\t60\t2S8M2D1I1M\t*\t0\t0\tNNNNNNNNNNNN\t*',
    # reverse strand, all aligned: span 6 from position 20
    'rev\t16\tchr\t20\t60\t6M\t*\t0\t0\tNNNNNN\t*',
    # half of the read soft-clipped: coverage 0.5
    'clipped\t0\tchr\t1\t60\t5M5S\t*\t0\t0\tNNNNNNNNNN\t*',
    'unmapped\t4\t*\t0\t0\t*\t*\t0\t0\tNNNN\t*',
    'secondary\t256\tchr\t3\t60\t4M\t*\t0\t0\tNNNN\t*',
    'supplementary_rev\t2064\tchr\t3\t60\t4M\t*\t0\t0\tNNNN\t*',
    'elsewhere\t0\tother_contig\t3\t60\t4M\t*\t0\t0\tNNNN\t*',
    'hardclip\t0\tchr\t50\t60\t3H4=1X3M\t*\t0\t0\tNNNNNNNN\t*']) + '\n'


@pytest.fixture(scope='module')
def cli():
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    return importlib.import_module('get_refs_from_sam')


def run(cli, tmp_path, capsys, *flags):
    (tmp_path / 'genome.fa').write_text('>chr some description\n' + GENOME[:30].lower() + '\n' + GENOME[30:] + '\n')
    (tmp_path / 'aln.sam').write_text(SAM)
    capsys.readouterr()
    n = cli.main(list(flags) + [str(tmp_path / 'genome.fa'), str(tmp_path / 'aln.sam')])
    lines = capsys.readouterr().out.strip().splitlines()
    got = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines), 2)}
    assert n == len(got)
    return got


def rc(s):
    return s.translate(str.maketrans('ACGTN', 'TGCAN'))[::-1]


def test_rules_on_hand_made_alignments(cli, tmp_path, capsys):
    # lower-case reference letters are outside ACGT and become N on loading (taiyaki/bio.py:59-72,
    # called with filter_ambig=False), as in the reference
    genome = 'N' * 30 + GENOME[30:]
    got = run(cli, tmp_path, capsys, '--min_coverage', '0.6')
    assert got == {'fwd': genome[4:15], 'rev': rc(genome[19:25]), 'hardclip': genome[49:57]}
    loose = run(cli, tmp_path, capsys, '--min_coverage', '0.5')
    assert sorted(loose) == ['clipped', 'fwd', 'hardclip', 'rev'] and loose['clipped'] == genome[0:5]
    padded = run(cli, tmp_path, capsys, '--pad', '3')
    assert padded == {'fwd': genome[1:18], 'rev': rc(genome[16:28]), 'hardclip': genome[46:57]}   # clipped at the end
    assert run(cli, tmp_path, capsys, '--reverse')['fwd'] == genome[4:15][::-1]
    assert run(cli, tmp_path, capsys, '--complement')['hardclip'] == rc(genome[49:57])[::-1]


def test_strand_list_output_file_and_compressed_input(cli, tmp_path, capsys):
    (tmp_path / 'strands.tsv').write_text('filename\tread_id\nx.fast5\trev\ny.fast5\tunmapped\n')
    assert run(cli, tmp_path, capsys, '--input_strand_list', str(tmp_path / 'strands.tsv')) == {
        'rev': rc(('N' * 30 + GENOME[30:])[19:25])}
    out = tmp_path / 'refs.fa'
    assert cli.main(['--output', str(out), str(tmp_path / 'genome.fa'), str(tmp_path / 'aln.sam'),
                     str(tmp_path / 'aln.sam')]) == 6                       # several inputs, in turn
    assert out.read_text().count('>') == 6
    with pytest.raises(SystemExit):
        cli.main(['--output', str(out), str(tmp_path / 'genome.fa'), str(tmp_path / 'aln.sam')])
    (tmp_path / 'aln.bam').write_bytes(b'\x1f\x8b\x08\x04' + bytes(20))
    with pytest.raises(SystemExit) as e:
        cli.main([str(tmp_path / 'genome.fa'), str(tmp_path / 'aln.bam')])
    assert 'samtools view' in str(e.value)


@needs_ref_files
def test_reference_sam_fixtures_against_the_genome(cli, capsys):
    """test/data/aligner_output/*.sam on test/data/genomic_reference.fasta: three of the five reads
    are aligned (two forward, one reverse); each record is the genome over the alignment's span --
    which is also where Guppy's alignment summary of the same reads puts it."""
    sams = sorted(os.path.join(REF_DATA, 'aligner_output', f)
                  for f in os.listdir(os.path.join(REF_DATA, 'aligner_output')) if f.endswith('.sam'))
    capsys.readouterr()
    assert cli.main([os.path.join(REF_DATA, 'genomic_reference.fasta')] + sams) == 3
    lines = capsys.readouterr().out.strip().splitlines()
    got = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines), 2)}
    genome = ''.join(line.strip() for line in open(os.path.join(REF_DATA, 'genomic_reference.fasta'))
                     if not line.startswith('>')).upper()
    summary = {}
    for line in open(os.path.join(REF_DATA, 'aligner_output', 'alignment_summary.txt')).read().splitlines()[1:]:
        f = line.split('\t')
        summary[f[0]] = (f[1], int(f[2]), int(f[3]))
    assert sorted(got) == sorted(k for k, v in summary.items() if v[0] != '*')
    for rid, seq in got.items():
        contig, start, end = summary[rid]           # 1-based start of the SAM record, span = end - start
        assert len(seq) == end - start
        want = genome[start - 1:end - 1]
        assert seq == (rc(want) if contig.endswith('_rc') else want)
    # the same reads under a stricter coverage: read 0f776a08 has 821 of its 2848 bases clipped
    assert cli.main(['--min_coverage', '0.8', os.path.join(REF_DATA, 'genomic_reference.fasta')] + sams) == 2
