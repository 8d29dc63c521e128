"""In-memory mapped reads and training chunks -- the part of
taiyaki/signal_mapping.py the training loop touches (SignalMapping :15-560,
Chunk :563-716), plus a synthetic r9.4.1-like read generator:
h5py is not available in this image, so `bin/train_flipflop.py` is fed these
objects through the same `chunk_selection.sample_chunks` surface."""
import numpy as np


class Chunk:
    """Signal + reference chunk with dwell filters (signal_mapping.py:563-716)."""
    _tiny = 0.00000001
    rej_str_pass = 'pass'
    rej_str_empty_seq = 'emptysequence'
    rej_str_empty_sig = 'emptysignal'
    rej_str_short = 'tooshort'
    rej_str_null_map = 'nullmapping'
    rej_str_path_buffer = 'pathbuffer'
    rej_str_mean_dwl = 'meandwell'
    rej_str_max_dwl = 'maxdwell'
    valid_rej_strs = set((rej_str_pass, rej_str_empty_seq, rej_str_empty_sig, rej_str_short,
                          rej_str_null_map, rej_str_path_buffer, rej_str_mean_dwl,
                          rej_str_max_dwl))

    def __init__(self, read_id, current=None, sequence=None, max_dwell=None,
                 start_sample=None, reject_reason=None):
        self.current = current
        self.sequence = sequence
        self.max_dwell = max_dwell
        self.start_sample = start_sample
        self.read_id = read_id
        self.reject_reason = self.rej_str_pass if reject_reason is None else reject_reason
        assert self.reject_reason in self.valid_rej_strs

    @property
    def accepted(self):
        return self.reject_reason == self.rej_str_pass

    @property
    def mean_dwell(self):
        return len(self.current) / (len(self.sequence) + self._tiny)

    @property
    def seq_len(self):
        return len(self.sequence) if self.sequence is not None else 0

    @property
    def sig_len(self):
        return len(self.current) if self.current is not None else 0

    def apply_filters(self, filter_params):
        if not self.accepted or filter_params.median_meandwell is None or \
           filter_params.mad_meandwell is None or filter_params.model_stride is None or \
           filter_params.path_buffer is None:
            return
        if self.sig_len / (self.seq_len * filter_params.model_stride) <= \
           filter_params.path_buffer:
            self.reject_reason = self.rej_str_path_buffer
            return
        if abs(self.mean_dwell - filter_params.median_meandwell) > \
           filter_params.filter_mean_dwell * filter_params.mad_meandwell:
            self.reject_reason = self.rej_str_mean_dwl
            return
        if self.max_dwell > filter_params.filter_max_dwell * filter_params.median_meandwell:
            self.reject_reason = self.rej_str_max_dwl


class SignalMapping:
    """A signal, a reference and the mapping between them (signal_mapping.py:15-560)."""

    def __init__(self, Dacs, Ref_to_signal, Reference, read_id='', shift_frompA=0.0,
                 scale_frompA=1.0, range=1.0, offset=0.0, digitisation=1.0):
        self.Dacs = np.asarray(Dacs, dtype=np.int16)
        self.Ref_to_signal = np.asarray(Ref_to_signal, dtype=np.int32)
        self.Reference = np.asarray(Reference, dtype=np.int16)
        self.read_id = read_id
        self.shift_frompA = float(shift_frompA)
        self.scale_frompA = float(scale_frompA)
        self.range = float(range)
        self.offset = float(offset)
        self.digitisation = float(digitisation)

    @property
    def siglen(self):
        return len(self.Dacs)

    @property
    def reflen(self):
        return len(self.Reference)

    @staticmethod
    def get_integer_reference(string_reference, alphabet):
        """Reference string -> int16 labels (signal_mapping.py:191-208)."""
        return np.array([alphabet.index(i) for i in string_reference], dtype=np.int16)

    @staticmethod
    def get_reftosignal(signalpos_to_refpos, reflen, siglen):
        """Reference position -> first signal sample, from a per-sample vector of reference
        positions with -1 for unmapped samples (signal_mapping.py:211-263): length reflen + 1;
        leading -1s for an unmapped start, trailing siglen + 1 for an unmapped end."""
        rts_dt = np.int32
        signalpos_to_refpos = np.asarray(signalpos_to_refpos)
        valid_idxs = np.where(signalpos_to_refpos != -1)[0].astype(rts_dt)
        if len(valid_idxs) == 0:
            return -1 * np.ones(reflen + 1, dtype=rts_dt)
        valid_refpos = signalpos_to_refpos[valid_idxs]
        move_pos = np.concatenate([[1, ], np.diff(valid_refpos)])
        ref_to_sig = np.repeat(valid_idxs, move_pos)
        ref_to_sig = np.concatenate([ref_to_sig, np.array([valid_idxs[-1] + 1, ], dtype=rts_dt)])
        if valid_refpos[0] > 0:
            ref_to_sig = np.concatenate([-1 * np.ones(valid_refpos[0], dtype=rts_dt), ref_to_sig])
        if reflen + 1 > len(ref_to_sig):
            ref_to_sig = np.append(ref_to_sig, (siglen + 1) * np.ones(
                reflen + 1 - len(ref_to_sig), dtype=rts_dt))
        return ref_to_sig

    @classmethod
    def from_remapping_path(cls, sigtoref_downsampled, reference, stride, untrimmed_dacs,
                            signalstart=0, **read_attrs):
        """Mapping from a remapping path over network blocks (signal_mapping.py:265-320):
        path element n belongs to sample stride * n - 1 + signalstart of the untrimmed
        signal.  The reference passes a `Signal` object; here its two fields used
        (`untrimmed_dacs`, `signalstart`) and the scaling attributes are passed directly."""
        rts_dt = np.int32
        sigtoref_downsampled = np.asarray(sigtoref_downsampled)
        fullsigtoref = np.full(len(untrimmed_dacs), -1, dtype=rts_dt)
        siglocs = np.arange(len(sigtoref_downsampled), dtype=rts_dt) * stride - 1 + signalstart
        f = np.logical_and(np.greater_equal(siglocs, 0), np.less(siglocs, len(fullsigtoref)))
        fullsigtoref[siglocs[f]] = sigtoref_downsampled[f]
        ref_to_sig = cls.get_reftosignal(fullsigtoref, reference.shape[0], len(untrimmed_dacs))
        return cls(untrimmed_dacs, ref_to_sig, reference, **read_attrs)

    def get_read_dictionary(self):
        """Fields as the mapped-signal writers take them (signal_mapping.py:322-345)."""
        return {'Dacs': self.Dacs, 'Ref_to_signal': self.Ref_to_signal, 'Reference': self.Reference,
                'read_id': self.read_id, 'shift_frompA': self.shift_frompA,
                'scale_frompA': self.scale_frompA, 'range': self.range, 'offset': self.offset,
                'digitisation': self.digitisation}

    def get_mapped_dacs_region(self):
        """(first, last) mapped sample (signal_mapping.py:208-225); cached: the arrays of a
        read do not change while it is being sampled from."""
        region = getattr(self, '_mapped_region', None)
        if region is None:
            valid = self.Ref_to_signal[np.logical_and(self.Ref_to_signal >= 0,
                                                      self.Ref_to_signal <= self.siglen)]
            region = (0, 0) if len(valid) == 0 else (int(valid[0]), int(valid[-1]))
            self._mapped_region = region
        return region

    def get_reference_locations(self, signal_location_vector):
        if isinstance(signal_location_vector, tuple):
            signal_location_vector = np.array(signal_location_vector)
        start, end = self.get_mapped_dacs_region()
        if any(signal_location_vector < start):
            raise IndexError('Signal location before mapped region requested.')
        if any(signal_location_vector > end):
            raise IndexError('Signal location after mapped region requested.')
        seq_start = np.searchsorted(self.Ref_to_signal, signal_location_vector[0], 'right') - 1
        seq_end = np.searchsorted(self.Ref_to_signal, signal_location_vector[1], 'left')
        return np.array([seq_start, seq_end])

    def get_current(self, region=None, standardize=True):
        dacs = self.Dacs if region is None else self.Dacs[region[0]:region[1]]
        current = (dacs + self.offset) * self.range / self.digitisation
        if standardize:
            current = (current - self.shift_frompA) / self.scale_frompA
        return current

    def _get_chunk(self, dacs_region, ref_region, standardize=True):
        if ref_region[1] == ref_region[0]:
            return Chunk(self.read_id, reject_reason=Chunk.rej_str_empty_seq)
        elif dacs_region[1] == dacs_region[0]:
            return Chunk(self.read_id, reject_reason=Chunk.rej_str_empty_sig)
        current = self.get_current(dacs_region, standardize)
        reference = self.Reference[ref_region[0]:ref_region[1]]
        dwells = np.diff(self.Ref_to_signal[ref_region[0]:ref_region[1]])
        maxdwell = np.max(dwells) if len(dwells) > 0 else 1
        return Chunk(self.read_id, current, reference, maxdwell, dacs_region[0])

    def get_chunk_with_sample_length(self, chunk_len, start_sample=None, standardize=True):
        region = self.get_mapped_dacs_region()
        spare_length = region[1] - region[0] - chunk_len
        if spare_length <= 0 or (start_sample is not None and start_sample >= spare_length):
            return Chunk(self.read_id, reject_reason=Chunk.rej_str_short)
        if start_sample is None:
            dacstart = np.random.randint(spare_length) + region[0]
        else:
            dacstart = start_sample + region[0]
        dacs_region = dacstart, chunk_len + dacstart
        try:
            ref_region = self.get_reference_locations(dacs_region)
        except IndexError:
            return Chunk(self.read_id, reject_reason=Chunk.rej_str_null_map)
        return self._get_chunk(dacs_region, ref_region, standardize)


def synthetic_reads(nreads, seed=7, min_len=40000, max_len=80000, mean_dwell=9.0,
                    nbase=4, kmer=6, mod_fraction=0.0, noise=0.3):
    """r9.4.1-like synthetic mapped reads: bases U{0..nbase-1}; dwell per base
    1 + Poisson(mean_dwell - 1); signal = level of the k-mer (table ~ N(0,1))
    + N(0, noise^2), quantised to int16 DACs with digitisation 1000.  With
    mod_fraction > 0 the alphabet is ACGT + 5mC ('Z' = label 4): each C is
    modified with that probability."""
    rng = np.random.RandomState(seed)
    table = rng.standard_normal(nbase ** kmer).astype(np.float32)
    reads = []
    for i in range(nreads):
        nsamp = rng.randint(min_len, max_len)
        nb = int(nsamp / mean_dwell)
        bases = rng.randint(0, nbase, size=nb)
        dwell = 1 + rng.poisson(mean_dwell - 1.0, size=nb)
        r2s = np.concatenate([[0], np.cumsum(dwell)]).astype(np.int32)
        idx = np.zeros(nb, dtype=np.int64)
        for k in range(kmer):
            sh = np.roll(bases, -k + kmer // 2)
            idx = idx * nbase + sh
        level = table[idx]
        sig = np.repeat(level, dwell) + noise * rng.standard_normal(r2s[-1]).astype(np.float32)
        dacs = np.clip(np.round(sig * 1000.0 / 4.0), -32000, 32000).astype(np.int16)
        ref = bases.astype(np.int16)
        if mod_fraction > 0:
            ref = ref.copy()
            ref[(bases == 1) & (rng.uniform(size=nb) < mod_fraction)] = nbase
        # Ref_to_signal has one entry per base plus the end position
        reads.append(SignalMapping(dacs, r2s, np.concatenate([ref, ref[-1:]]),
                                   read_id='synthetic_%06d' % i, shift_frompA=0.0,
                                   scale_frompA=1.0, range=4.0, offset=0.0, digitisation=1000.0))
    return reads
