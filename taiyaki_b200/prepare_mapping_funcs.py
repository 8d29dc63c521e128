"""Remapping reads to their references with a flip-flop model -- the flow of
taiyaki/prepare_mapping_funcs.py:24-109 (`oneread_remap`) behind bin/prepare_mapped_reads.py,
for raw signals passed in (bin/prepare_mapped_reads.py reads them from fast5 files through
fast5utils / signal.Signal): trim and standardise with the per-read parameters, run the network over the whole read, align the reference to the
transition scores (csrc/remap.cu), turn the block path into Ref_to_signal.

`remap_reads` does this for a list of reads with ONE remapping launch for all of them (one
CTA per read) instead of one worker process per read."""
import enum
import sys

import numpy as np
import torch

from . import flipflop_remap, helpers
from .signal_mapping import SignalMapping


class RemapResult(enum.Enum):
    """Possible results of remapping a read (prepare_mapping_funcs.py:14-21)."""
    SUCCESS = 'Success!'
    READ_ID_INFO_NOT_FOUND = 'No information for read id found in file.'
    NO_REF_FOUND = 'No fasta reference found.'
    NO_PARAMS = 'No per-read params provided.'
    NETWORK_ERROR = 'Failure applying basecall network to remap read.'
    REF_TOO_LONG = 'Reference exceeded maximum allowed read length.'


def get_per_read_params_dict_from_tsv(input_file):
    """UUID -> {trim_start, trim_end, shift, scale} from a per-read parameter .tsv
    (prepare_mapping_funcs.py:148-177)."""
    out = {}
    with open(input_file) as fh:
        header = fh.readline().rstrip('\n').split('\t')
        col = {name: header.index(name) for name in ('UUID', 'trim_start', 'trim_end', 'shift', 'scale')}
        for line in fh:
            f = line.rstrip('\n').split('\t')
            try:
                out[f[col['UUID']]] = {'trim_start': int(f[col['trim_start']]),
                                       'trim_end': int(f[col['trim_end']]),
                                       'shift': float(f[col['shift']]), 'scale': float(f[col['scale']])}
            except Exception:
                sys.stderr.write('Warning: ignoring incorrect line {} in {}\n'.format(f, input_file))
    return out


def fasta_file_to_dict(fasta_file_name, filter_ambig=True, flatten_ambig=True, alphabet='ACGT'):
    """Record id (up to the first blank) -> sequence (taiyaki/bio.py:43-81).  Empty records are
    dropped; with `filter_ambig` so are records holding a character outside `alphabet` (a count of
    them goes to stderr); otherwise, with `flatten_ambig`, such characters become N."""
    import re
    notbase = re.compile('[^{}]'.format(re.escape(alphabet)))
    references, skipped = {}, 0

    def keep(name, parts):
        nonlocal skipped
        seq = ''.join(parts)
        if name is None or len(seq) == 0:
            return
        if filter_ambig and notbase.search(seq) is not None:
            skipped += 1
            return
        references[name] = notbase.sub('N', seq) if flatten_ambig else seq
    name, parts = None, []
    with open(fasta_file_name) as fh:
        for line in fh:
            line = line.strip()
            if line.startswith('>'):
                keep(name, parts)
                fields = line[1:].split()
                name, parts = (fields[0] if fields else ''), []
            elif line:
                parts.append(line)
    keep(name, parts)
    if skipped > 0:
        sys.stderr.write('* {} reference seqeunces contain ambiguous bases not found in the provided '
                         'alphabet and will be skipped.\n'.format(skipped))
    return references


def generate_output_from_results(results, output, alphabet_info, verbose=True, batch_format=True):
    """Write the successful remappings to a mapped-signal file and report the failures by
    category (prepare_mapping_funcs.py:112-145)."""
    from collections import defaultdict
    from .mapped_signal_files import MappedSignalWriter
    err_types = defaultdict(int)
    count = 0
    with MappedSignalWriter(output, alphabet_info, batch_format) as msw:
        for resultdict, mesg in results:
            if resultdict is None:
                err_types[mesg] += 1
            else:
                count += 1
                msw.write_read(resultdict)
    sys.stderr.write('* {} reads mapped successfully\n'.format(count))
    for result, n_errs in err_types.items():
        sys.stderr.write('* {} reads failed to produce remapping results due to: {}\n'.format(
            n_errs, getattr(result, 'value', result)))
    return count, dict(err_types)


def trim_bounds(nsample, trim_start, trim_end):
    """(signalstart, signalend_exc) of taiyaki/signal.py:77-95 (set_trim_absolute)."""
    if trim_start < 0 or trim_end < 0:
        raise Exception("Can't trim a negative amount off the end of a signal vector.")
    if trim_start + trim_end >= nsample:
        trim_start, trim_end = 0, 0
    return trim_start, nsample - trim_end


def remap_reads(reads, model, per_read_params_dict, alphabet_info, max_read_length=None,
                localpen=0.0, model_stride=None):
    """`reads`: list of dicts with read_id, dacs (untrimmed int16; None when the read could not be
    loaded), offset, range, digitisation and ref (reference string, possibly with modified bases, or None).
    Returns [(read dictionary or None, RemapResult)] in the same order -- what
    `oneread_remap` returns per read."""
    device = helpers.get_model_device(model)
    if model_stride is None:
        model_stride = helpers.guess_model_stride(model)
    results = [None] * len(reads)
    todo = []
    with torch.no_grad():
        for i, read in enumerate(reads):
            read_ref = read.get('ref')
            if read_ref is None:
                results[i] = (None, RemapResult.NO_REF_FOUND)
                continue
            if max_read_length is not None and len(read_ref) > max_read_length:
                results[i] = (None, RemapResult.REF_TOO_LONG)
                continue
            params = per_read_params_dict.get(read['read_id'])
            if params is None:
                results[i] = (None, RemapResult.NO_PARAMS)
                continue
            if read.get('dacs') is None:          # the file did not yield this read's signal
                results[i] = (None, RemapResult.READ_ID_INFO_NOT_FOUND)
                continue
            dacs = np.asarray(read['dacs'])
            start, end = trim_bounds(len(dacs), int(params['trim_start']), int(params['trim_end']))
            try:
                current = (dacs[start:end] + read['offset']) * read['range'] / read['digitisation']
                standardized = ((current - params['shift']) / params['scale']).astype(np.float32)
                signal = torch.as_tensor(standardized[:, None, None]).to(device)
                transweights = model(signal)[:, 0, :]
            except Exception:
                results[i] = (None, RemapResult.NETWORK_ERROR)
                continue
            can_read_ref = alphabet_info.collapse_sequence(read_ref)
            todo.append((i, transweights, can_read_ref, start, params))
        if todo:
            aligned = flipflop_remap.flipflop_remap_batch(
                [t[1] for t in todo], [t[2] for t in todo], alphabet=alphabet_info.can_bases,
                localpen=localpen)
            for (i, _, _, start, params), (_, path) in zip(todo, aligned):
                read = reads[i]
                # a bad reference or mapping fails THIS read, not the batch (the reference's
                # oneread_remap returns (None, message) per read, prepare_mapping_funcs.py:77-110,
                # and SignalMapping.check() keeps malformed mappings out of the training file)
                try:
                    int_ref = SignalMapping.get_integer_reference(read['ref'], alphabet_info.alphabet)
                    sig_mapping = SignalMapping.from_remapping_path(
                        path, int_ref, model_stride, np.asarray(read['dacs']), start,
                        read_id=read['read_id'], shift_frompA=params['shift'],
                        scale_frompA=params['scale'], range=read['range'], offset=read['offset'],
                        digitisation=read['digitisation'])
                    from .mapped_signal_files import check_read
                    problem = check_read(sig_mapping)
                except Exception as e:
                    problem = 'remapping failed: %s' % e
                if problem != 'pass':
                    results[i] = (None, problem)
                    continue
                results[i] = (sig_mapping.get_read_dictionary(), RemapResult.SUCCESS)
    return results
