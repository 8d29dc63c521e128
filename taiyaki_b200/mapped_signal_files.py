"""Readers of mapped-signal files -- the reading half of taiyaki/mapped_signal_files.py
(AbstractMappedSignalReader :26-203, PerReadHDF5Reader :255-341, BatchHDF5Reader :427-559,
HDF5Reader / MappedSignalReader :681-731) on the plain-Python HDF5
decoder of hdf5_min.py, because this image has no h5py.  Same class and method names and
the same file layouts (docs/FILE_FORMATS.md:40-88):

  per read : /Reads/<read_id>/{Dacs, Ref_to_signal, Reference} + attributes
  batched  : /Batches/Batch_<n>/{Dacs, Dacs_lengths, ..., read_id, shift_frompA, ...}
  both     : root attributes version, alphabet, collapse_alphabet, mod_long_names

Reads come back as `signal_mapping.SignalMapping`, what `chunk_selection` and the device
read store consume.  `BatchHDF5Writer` writes the batched layout through hdf5_min_write.py."""
import inspect

import numpy as np

from . import hdf5_min
from .alphabet import AlphabetInfo
from .signal_mapping import SignalMapping

_version = 8
READS_ROOT_TEXT = 'Reads'
BATCH_ROOT_TEXT = 'Batches'
BATCH_TMPLT = 'Batch_{}'
BATCH_LENGTH_SUFFIX = '_lengths'
_ARRAY_KEYS = ('Dacs', 'Ref_to_signal', 'Reference')
_READ_KWARGS = frozenset(inspect.signature(SignalMapping.__init__).parameters) - {'self'}


def _signal_mapping(d):
    """SignalMapping from a read dictionary; optional fields this path does not use
    (mapping_score, mapping_method) are dropped."""
    return SignalMapping(**{k: v for k, v in d.items() if k in _READ_KWARGS})


def check_read(read):
    """Integrity checks of SignalMapping.check (signal_mapping.py:87-116)."""
    msg = ''
    maplen = len(read.Ref_to_signal)
    if read.reflen + 1 != maplen:
        msg += ('Length of Ref_to_signal ({}) should be 1 + length of Reference ({})\n').format(
            maplen, read.reflen)
    if np.min(read.Ref_to_signal) < -1 or np.max(read.Ref_to_signal) > len(read.Dacs) + 1:
        msg += 'Range of locations in mapping exceeds length of Dacs\n'
    if np.any(np.diff(read.Ref_to_signal) < 0):
        msg += 'Mapping does not increase monotonically\n'
    return 'pass' if len(msg) == 0 else msg


class AbstractMappedSignalReader:
    """Shared behaviour (mapped_signal_files.py:26-203)."""
    pass_str = 'pass'

    def __init__(self, filename):
        self.hdf5 = hdf5_min.File(filename)
        assert self.version == _version, (
            'Incorrect file version, got {} expected {}').format(self.version, _version)

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()

    def close(self):
        self.hdf5.close()

    @property
    def version(self):
        return int(self.hdf5.attrs['version'])

    def get_alphabet_information(self):
        mod_long_names = self.hdf5.attrs['mod_long_names'].splitlines()
        return AlphabetInfo(self.hdf5.attrs['alphabet'], self.hdf5.attrs['collapse_alphabet'],
                            mod_long_names)

    def reads(self, read_ids=None):
        if read_ids is None:
            yield from self
        else:
            yield from self._some_reads(read_ids)

    def check(self, limit_report_lines=100):
        return_string = ''
        file_is_empty = True
        for read in self:
            file_is_empty = False
            if return_string.count('\n') >= limit_report_lines:
                return_string += ('----------Number of lines in error report limited to ' +
                                  str(limit_report_lines) + '\n')
                break
            read_check = check_read(read)
            if read_check != self.pass_str:
                return_string += 'Read ' + read.read_id + ':\n' + read_check
        if file_is_empty:
            return_string += 'No reads in file\n'
        return self.pass_str if len(return_string) == 0 else return_string


class PerReadHDF5Reader(AbstractMappedSignalReader):
    """One group per read (mapped_signal_files.py:255-341)."""

    def __init__(self, filename, load_in_mem=False):
        super().__init__(filename)

    def __iter__(self):
        self.reads_iter = iter(self.hdf5[READS_ROOT_TEXT].keys())
        return self

    def __next__(self):
        return self.get_read(next(self.reads_iter))

    def _some_reads(self, read_ids):
        present = self.get_read_ids()
        wanted = set(read_ids)
        for read_id in present:
            if read_id in wanted:
                yield self.get_read(read_id)

    def get_read(self, read_id):
        group = self.hdf5[READS_ROOT_TEXT + '/' + read_id]
        d = {k: group[k][()] for k in group.keys()}
        d.update(group.attrs)
        return _signal_mapping(d)

    def get_read_ids(self):
        if 'read_ids' in self.hdf5.root:
            return [str(r) for r in self.hdf5['read_ids'][()].tolist()]
        try:
            return list(self.hdf5[READS_ROOT_TEXT].keys())
        except Exception:
            return []


class BatchHDF5Reader(AbstractMappedSignalReader):
    """Reads stored as concatenated arrays per batch (mapped_signal_files.py:427-559)."""

    def __init__(self, filename):
        super().__init__(filename)
        self.batch_names = list(self.hdf5[BATCH_ROOT_TEXT].keys())
        self.read_id_to_batch_str = {}
        for batch_name in self.batch_names:
            for read_id in self.hdf5[BATCH_ROOT_TEXT + '/' + batch_name + '/read_id'][()]:
                self.read_id_to_batch_str[str(read_id)] = batch_name
        self._cache = (None, None)

    def __iter__(self):
        self._iter = (read for batch_name in self.batch_names
                      for read in self._load_reads_batch(batch_name).values())
        return self

    def __next__(self):
        return next(self._iter)

    def _some_reads(self, read_ids):
        """The wanted reads batch by batch (each batch decoded once), in file order within a batch."""
        wanted = set(read_ids)
        by_batch = {}
        for read_id, name in self.read_id_to_batch_str.items():
            if read_id in wanted:
                by_batch.setdefault(name, []).append(read_id)
        for batch_name in self.batch_names:
            if batch_name in by_batch:
                batch = self._load_reads_batch(batch_name)
                for read_id in by_batch[batch_name]:
                    yield batch[read_id]

    def _load_reads_batch(self, batch_name):
        if batch_name not in self.batch_names:
            raise RuntimeError('Invalid batch name requested: {}'.format(batch_name))
        if self._cache[0] == batch_name:
            return self._cache[1]
        group = self.hdf5[BATCH_ROOT_TEXT + '/' + batch_name]
        keys = [k for k in group.keys() if not k.endswith(BATCH_LENGTH_SUFFIX)]
        columns = []
        for k in keys:
            val = group[k][()]
            if k in _ARRAY_KEYS or (k + BATCH_LENGTH_SUFFIX) in group:
                val = np.split(val, np.cumsum(group[k + BATCH_LENGTH_SUFFIX][()][:-1]))
            columns.append(val)
        parsed = {}
        for values in zip(*columns):
            d = dict(zip(keys, values))
            d['read_id'] = str(d['read_id'])
            parsed[d['read_id']] = _signal_mapping(d)
        self._cache = (batch_name, parsed)
        return parsed

    def get_read(self, read_id):
        return self._load_reads_batch(self.read_id_to_batch_str[read_id])[read_id]

    def get_read_ids(self):
        return list(self.read_id_to_batch_str.keys())


def HDF5Reader(filename, load_in_mem=False):
    """Per-read or batched reader, by the groups present (mapped_signal_files.py:681-705)."""
    probe = hdf5_min.File(filename)
    is_batch = BATCH_ROOT_TEXT in probe.root
    probe.close()
    if is_batch:
        return BatchHDF5Reader(filename)
    return PerReadHDF5Reader(filename, load_in_mem)


class BatchHDF5Writer:
    """Batched mapped-signal file writer with the reference's interface
    (mapped_signal_files.py:562-679): `write_read(readdict)` with the dictionaries of
    `SignalMapping.get_read_dictionary`, `close()`, context manager.  Reads are collected
    and laid out by hdf5_min_write.py when the file is closed (see its status note)."""

    def __init__(self, filename, alphabet_info, batch_size=25000):
        self.filename = filename
        self.alphabet_info = alphabet_info
        self.batch_size = batch_size
        self.read_ids = []
        self._reads = []

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()

    def write_read(self, readdict):
        self.read_ids.append(readdict['read_id'])
        self._reads.append(readdict)

    def close(self):
        from .hdf5_min_write import write_batched_mapped_signal_file
        ai = self.alphabet_info
        longest = max([len(r['Dacs']) for r in self._reads] + [1])
        write_batched_mapped_signal_file(
            self.filename, self._reads, batch_size=self.batch_size,
            chunk=int(min(max(longest, 1 << 16), 1 << 20)),
            alphabet=(ai.alphabet, ai.collapse_alphabet, '\n'.join(ai.mod_long_names or [])))


def HDF5Writer(filename, alphabet_info, batch_format=True):
    """mapped_signal_files.py:708-726; only the batched format is written here."""
    if not batch_format:
        raise NotImplementedError('the per-read format (libver v108 headers) is not written here')
    return BatchHDF5Writer(filename, alphabet_info)


# module-level swap points, as in the reference (mapped_signal_files.py:729-731)
MappedSignalReader = HDF5Reader
MappedSignalWriter = HDF5Writer
