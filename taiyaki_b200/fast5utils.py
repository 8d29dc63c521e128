"""Raw reads from ONT fast5 files for the callers either side of the hot path
(bin/prepare_mapped_reads.py, bin/generate_per_read_params.py, bin/basecall.py), behind the
names of taiyaki/fast5utils.py (:16-180 iteration over files / strand lists, :187-290 channel
and read attributes).

The reference reaches these files through ont_fast5_api + h5py, neither of which is in this
image; here the same two layouts are decoded by the plain-Python HDF5 reader of this package
(hdf5_min.py):

  single-read file   /Raw/Reads/Read_<n>  (attributes read_id, read_number, ...; dataset Signal)
                     /UniqueGlobalKey/{channel_id, context_tags, tracking_id}
  multi-read file    /read_<uuid>/Raw     (the same attributes; dataset Signal)
                     /read_<uuid>/{channel_id, context_tags, tracking_id}

`Fast5File.get_read_ids()` / `.get_read(read_id)` and the read object's `.handle`,
`.global_key` and `.get_raw_data()` are the part of ont_fast5_api's interface the reference
uses, so `get_channel_info(read)` etc. read as they do there.  Signals stored with the VBZ
filter (HDF5 filter 32020, MinKNOW >= 19.12) are not decoded: hdf5_min raises
`Hdf5FormatError` naming the filter and the read is reported as unreadable, like any other
read the reference fails to load.
"""
import os
import sys

from . import hdf5_min

SINGLE_READ, MULTI_READ = 'single-read', 'multi-read'


def _text(v):
    return v.decode('utf-8') if isinstance(v, bytes) else str(v)


class Fast5Read:
    """One read of a fast5 file: `handle` is the group its keys are relative to (the file for a
    single-read file, /read_<uuid> for a multi-read one), `global_key` the prefix of the
    channel / context / tracking groups below it."""

    def __init__(self, handle, global_key, raw_group, read_id):
        self.handle, self.global_key = handle, global_key
        self.raw_dataset_group_name = raw_group
        self.read_id = read_id

    def get_read_id(self):
        return self.read_id

    def get_raw_data(self):
        """int16 DAC samples of the whole read."""
        return self.handle[self.raw_dataset_group_name + '/Signal'].read()


class Fast5File:
    def __init__(self, filename, mode='r'):
        if mode != 'r':
            raise ValueError('fast5 files are read-only here')
        self.filename = filename
        self.handle = hdf5_min.File(filename)
        self.file_type = self._file_type()

    def _file_type(self):
        declared = self.handle.attrs.get('file_type')
        if declared is not None and _text(declared) in (SINGLE_READ, MULTI_READ):
            return _text(declared)
        keys = list(self.handle.keys())
        if any(k.startswith('read_') for k in keys):
            return MULTI_READ
        if len(keys) == 0 or 'UniqueGlobalKey' in keys:
            return SINGLE_READ
        raise TypeError('{} is neither a single- nor a multi-read fast5 file'.format(self.filename))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        self.handle.close()

    def _single_reads(self):
        """Read_<n> group name -> read id, in a single-read file."""
        try:
            reads = self.handle['Raw/Reads']
        except KeyError:
            return {}
        return {k: _text(reads[k].attrs['read_id']) for k in reads.keys()}

    def get_read_ids(self):
        if self.file_type == MULTI_READ:
            return [k[len('read_'):] for k in self.handle.keys() if k.startswith('read_')]
        return list(self._single_reads().values())

    def get_read(self, read_id):
        if self.file_type == MULTI_READ:
            try:
                return Fast5Read(self.handle['read_' + read_id], '', 'Raw', read_id)
            except KeyError:
                raise KeyError('read {} not in {}'.format(read_id, self.filename))
        for group, rid in self._single_reads().items():
            if rid == read_id:
                return Fast5Read(self.handle, 'UniqueGlobalKey/', 'Raw/Reads/' + group, read_id)
        raise KeyError('read {} not in {}'.format(read_id, self.filename))


def get_fast5_file(filename, mode='r'):
    return Fast5File(filename, mode)


class ReadLoader:
    """get_read(filename, read_id) that keeps the file of the previous call open: consecutive reads
    of a multi-read file (4000 reads per file is usual) then share one parse of the file's root
    group instead of repeating it per read."""

    def __init__(self):
        self._name, self._file = None, None

    def get_read(self, filename, read_id):
        if filename != self._name:
            self.close()
            self._file, self._name = get_fast5_file(filename, 'r'), filename
        return self._file.get_read(read_id)

    def close(self):
        if self._file is not None:
            self._file.close()
        self._name, self._file = None, None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def get_fast5_file_list(path, recursive=False):
    """*.fast5 below `path` (a file is returned as it is); sorted, so that runs are repeatable."""
    if os.path.isfile(path):
        return [path]
    found = []
    for root, dirs, files in os.walk(path):
        found.extend(os.path.join(root, f) for f in files if f.endswith('.fast5'))
        if not recursive:
            break
    return sorted(found)


def _skipped(err):
    sys.stderr.write('Warning: An exception occured in fast5utils (skipped this read):\n{}\n'.format(err))


def iterate_file_read_pairs(filepaths, read_ids, limit=None, verbose=0):
    """(file, read id) rows of a strand list: those whose file exists and holds the read
    (fast5utils.py:16-49)."""
    nyielded = 0
    listed = (None, frozenset())          # read ids of the file of the previous row
    for filepath, read_id in zip(filepaths, read_ids):
        if not os.path.exists(filepath):
            sys.stderr.write('File {} does not exist, skipping\n'.format(filepath))
            continue
        if filepath != listed[0]:
            try:
                with get_fast5_file(filepath, 'r') as f5file:
                    listed = (filepath, frozenset(f5file.get_read_ids()))
            except Exception as e:
                _skipped(e)
                continue
        if read_id not in listed[1]:
            continue
        if verbose > 0:
            print('Reading', read_id, 'from', filepath)
        yield filepath, read_id
        nyielded += 1
        if limit is not None and nyielded >= limit:
            return


def iterate_files_reads_unpaired(filepaths, read_ids, limit=None, verbose=0):
    """Every read of every file, kept when `read_ids` is None or contains it
    (fast5utils.py:52-92)."""
    wanted = None if read_ids is None else frozenset(read_ids)
    nyielded = 0
    for filepath in filepaths:
        if not os.path.exists(filepath):
            sys.stderr.write('File {} does not exist, skipping\n'.format(filepath))
            continue
        try:
            with get_fast5_file(filepath, 'r') as f5file:
                ids = f5file.get_read_ids()
        except Exception as e:
            _skipped(e)
            continue
        for read_id in ids:
            if wanted is None or read_id in wanted:
                if verbose > 0:
                    print('Reading', read_id, 'from', filepath)
                yield filepath, read_id
                nyielded += 1
            elif verbose > 0:
                print('Skipping', read_id, 'from', filepath, ':not in read_id list')
            if limit is not None and nyielded >= limit:
                return


def _strand_list_columns(strand_list):
    """Header-named columns of a tab-separated strand list (also a sequencing summary)."""
    with open(strand_list) as fh:
        names = fh.readline().rstrip('\n').split('\t')
        rows = [line.rstrip('\n').split('\t') for line in fh if line.strip()]
    return {nm: [r[i] for r in rows if i < len(r)] for i, nm in enumerate(names)}


def iterate_fast5_reads(path, strand_list=None, limit=None, verbose=0, recursive=False):
    """(file path, read id) for the reads of a directory of fast5 files (single- or multi-read)
    or of one file (fast5utils.py:95-180).  With a strand list: a `read_id` column alone
    selects those reads from all files below `path`; a `filename` / `filename_fast5` column
    alone selects all reads of those files; both give (file, read) rows that are checked."""
    filepaths, read_ids = None, None
    if strand_list is not None:
        table = _strand_list_columns(strand_list)
        if verbose >= 2:
            print('Columns in strand list file:')
            print(tuple(table))
        if 'filename' in table:
            filepaths = table['filename']
        elif 'filename_fast5' in table:
            filepaths = table['filename_fast5']
        if 'read_id' in table:
            read_ids = [str(i) for i in table['read_id']]
        if filepaths is None and read_ids is None:
            raise Exception(
                "Strand list at {} has no column that can be used: (it should contain ('filename' or "
                "'filename_fast5') or 'read_id', or both a filename column and a read_id "
                "column)".format(strand_list))
        if filepaths is not None:       # the list holds file names; `path` supplies the directory
            filepaths = [os.path.join(path, x) for x in filepaths]
    if filepaths is not None and read_ids is not None:
        yield from iterate_file_read_pairs(filepaths, read_ids, limit, verbose)
        return
    if filepaths is None:
        filepaths = get_fast5_file_list(path, recursive=recursive)
    yield from iterate_files_reads_unpaired(filepaths, read_ids, limit, verbose)


def get_filename(read):
    """`filename` of the read's context tags (fast5utils.py:187-210)."""
    return read.handle[read.global_key + 'context_tags'].attrs['filename']


def get_channel_info(read):
    """digitisation, range, offset, sampling_rate, channel_number (fast5utils.py:213-238)."""
    return read.handle[read.global_key + 'channel_id'].attrs


def get_read_attributes(read):
    """start_time, duration, read_id, ...: on /Raw in a multi-read file, on the highest-numbered
    /Raw/Reads/Read_<n> in a single-read one (fast5utils.py:241-275)."""
    r = read.handle['Raw'].attrs
    if len(r) > 0:
        return r
    numbered_reads = list(read.handle['Raw/Reads'].keys())
    return read.handle['Raw/Reads/' + sorted(numbered_reads)[-1]].attrs
