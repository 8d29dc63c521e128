"""Test helper: fast5 files (single- and multi-read layout) written with the package's minimal
HDF5 writer from the raw reads stored in tests/golden/prepare_remap.npz, so that the GPU box
-- which has no /root/reference -- gets the reference's fixture reads as files of the layout
test/data/reads and test/data/multireads have (group names, attribute names and types as in
those files; contents from the golden file)."""
import os

import numpy as np

from taiyaki_b200.hdf5_min_write import Writer

HERE = os.path.dirname(os.path.abspath(__file__))
CHANNEL_KEYS = ('offset', 'range', 'digitisation', 'sampling_rate')
READ_ATTR_KEYS = (('start_time', np.uint64), ('duration', np.uint32), ('read_number', np.int32),
                  ('start_mux', np.uint8))


def golden():
    return np.load(os.path.join(HERE, 'golden', 'prepare_remap.npz'))


def _read_groups(w, g, rid, filename):
    """(Raw attributes, Signal dataset, channel_id / context_tags / tracking_id groups) of a read."""
    attrs = [(k, dt(v)) for (k, dt), v in zip(READ_ATTR_KEYS, g[rid + '_read_attrs'])]
    attrs.append(('read_id', np.bytes_(rid.encode())))
    signal = w.dataset(np.asarray(g[rid + '_dacs'], dtype=np.int16), 4096)
    channel = w.group({}, attrs=[(k, np.float64(v)) for k, v in zip(CHANNEL_KEYS, g[rid + '_channel'])] +
                      [('channel_number', np.bytes_(b'9'))])[0]
    context = w.group({}, attrs=[('filename', np.bytes_(filename.encode())),
                                 ('experiment_type', np.bytes_(b'genomic_dna'))])[0]
    tracking = w.group({}, attrs=[('device_id', np.bytes_(b'MN17205'))])[0]
    return attrs, signal, channel, context, tracking


def write_single_read_fast5(path, g, rid):
    w = Writer()
    attrs, signal, channel, context, tracking = _read_groups(w, g, rid, os.path.basename(path))
    read = w.group({'Signal': signal}, attrs=attrs)[0]
    reads = w.group({str(g[rid + '_read_group']): read})[0]
    raw = w.group({'Reads': reads})[0]
    unique = w.group({'channel_id': channel, 'context_tags': context, 'tracking_id': tracking})[0]
    root = w.group({'Raw': raw, 'UniqueGlobalKey': unique}, attrs=[('file_version', np.float64(2.0))])
    w.close(*root, path)


def write_multi_read_fast5(path, g, rids):
    w = Writer()
    top = {}
    for rid in rids:
        attrs, signal, channel, context, tracking = _read_groups(w, g, rid, os.path.basename(path))
        raw = w.group({'Signal': signal}, attrs=attrs)[0]
        top['read_' + rid] = w.group({'Raw': raw, 'channel_id': channel, 'context_tags': context,
                                      'tracking_id': tracking})[0]
    w.close(*w.group(top), path)


def write_inputs(folder, g, multi=False):
    """fast5 files, per-read parameter table and reference fasta of the golden reads below
    `folder`; returns (reads directory, tsv path, fasta path)."""
    rids = [str(r) for r in g['read_ids']]
    reads_dir = os.path.join(str(folder), 'reads')
    os.makedirs(reads_dir, exist_ok=True)
    if multi:
        write_multi_read_fast5(os.path.join(reads_dir, 'batch_0.fast5'), g, rids)
    else:
        for rid in rids:
            write_single_read_fast5(os.path.join(reads_dir, rid + '.fast5'), g, rid)
    tsv = os.path.join(str(folder), 'readparams.tsv')
    with open(tsv, 'w') as fh:
        fh.write('UUID\ttrim_start\ttrim_end\tshift\tscale\n')
        for rid in rids:
            t0, t1, shift, scale = g[rid + '_params']
            fh.write('{}\t{}\t{}\t{!r}\t{!r}\n'.format(rid, int(t0), int(t1), float(shift), float(scale)))
    fasta = os.path.join(str(folder), 'refs.fasta')
    with open(fasta, 'w') as fh:
        for rid in rids:
            if rid + '_reference' in g:
                fh.write('>{}\n{}\n'.format(rid, str(g[rid + '_reference'])))
    return reads_dir, tsv, fasta
