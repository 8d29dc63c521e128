"""Alphabet description consumed by the model-definition API
(taiyaki/alphabet.py:4-248, the attributes the flip-flop layers read)."""
import numpy as np


class AlphabetInfo(object):
    def __init__(self, alphabet, collapse_alphabet, mod_long_names=[], do_reorder=False):
        self.alphabet = alphabet
        self.collapse_alphabet = collapse_alphabet
        self.mod_long_names = mod_long_names
        try:
            self.alphabet = self.alphabet.decode()
            self.collapse_alphabet = self.collapse_alphabet.decode()
        except Exception:
            pass
        self.parse_alphabet_info()
        self.validate_alphabet()
        self.is_sorted = False
        if do_reorder:
            self.sort_alphabet()

    def parse_alphabet_info(self):
        self.translation_table = self.alphabet.maketrans(self.alphabet, self.collapse_alphabet)
        self.nbase = len(self.alphabet)
        self.can_bases_set = set(self.collapse_alphabet)
        self.mod_bases_set = set(self.alphabet).difference(self.can_bases_set)
        mod_bases = [b for b in self.alphabet if b in self.mod_bases_set]
        self.mod_name_conv = (None if self.mod_long_names is None else
                              dict(zip(mod_bases, self.mod_long_names)))
        self.ncan_base = len(self.can_bases_set)
        self.nmod_base = self.nbase - self.ncan_base
        self.add_ordered_info()

    def add_ordered_info(self):
        self.collapse_labels = np.array(
            [self.alphabet.find(cb) for cb in self.collapse_alphabet], dtype=np.int32)
        self.can_bases = ''.join([b for b in self.alphabet if b in self.can_bases_set])
        self.mod_bases = ''.join([b for b in self.alphabet if b in self.mod_bases_set])

    def sort_alphabet(self):
        self.collapse_alphabet, self.alphabet = map(
            lambda x: ''.join(x), zip(*sorted(zip(self.collapse_alphabet, self.alphabet))))
        if self.mod_long_names is not None:
            self.mod_long_names = [self.mod_name_conv[b] for b in self.alphabet
                                   if b in self.mod_bases_set]
        self.is_sorted = True
        self.add_ordered_info()

    def validate_alphabet(self):
        assert len(self.alphabet) == len(self.collapse_labels)
        assert len(set(self.collapse_alphabet).difference(self.alphabet)) == 0, (
            'All bases in collapse alphabet must occur within alphabet.')
        if self.nmod_base > 0:
            assert self.mod_long_names is not None
            assert self.nmod_base == len(self.mod_long_names)

    def collapse_sequence(self, sequence_with_mods):
        """Modified bases -> their canonical bases (alphabet.py:120-124)."""
        return sequence_with_mods.translate(self.translation_table)

    def _label_counts(self, read_data, nreads):
        """Occurrences of each label in the references of `nreads` reads drawn without
        replacement from `read_data` (objects with .Reference, or read dictionaries)."""
        picked = np.random.choice(len(read_data), min(nreads, len(read_data)), replace=False)
        refs = [read_data[i] for i in picked]
        refs = [r['Reference'] if isinstance(r, dict) else r.Reference for r in refs]
        counts = np.bincount(np.concatenate(refs).astype(np.int64))
        if len(counts) < self.nbase or not counts.all():
            # a label that never occurs has no frequency ratio (alphabet.py:58-59, :89-90)
            raise NotImplementedError
        return counts

    def _alternatives(self, can_label):
        """Labels collapsing onto `can_label`, without the first of them -- the canonical base
        itself in an alphabet that lists canonical bases before their modifications."""
        return np.flatnonzero(self.collapse_labels == can_label)[1:]

    def compute_mod_inv_freq_weights(self, read_data, N):
        """canonical count / modified count per modified base, 1 per canonical base, in the
        output order of the cat-mod layer (alphabet.py:35-66)."""
        counts = self._label_counts(read_data, N)
        weights = []
        for can_label in range(self.ncan_base):
            weights.append(1.0)
            weights.extend(counts[can_label] / counts[m] for m in self._alternatives(can_label))
        return np.array(weights, dtype=np.float32)

    def compute_log_odds_weights(self, read_data, N):
        """Prior odds of the modified-base categories from `N` sampled reads, in the output order
        of the cat-mod layer: (sum of modified counts) / canonical count for each canonical base,
        then canonical count / modified count for each of its modifications
        (alphabet.py:68-100; used by --mod_prior_factor, train_flipflop.py:312-326)."""
        counts = self._label_counts(read_data, N)
        weights = []
        for base in self.can_bases:
            can_label = self.alphabet.index(base)
            alts = self._alternatives(can_label)
            weights.append(counts[alts].sum() / counts[can_label])
            weights.extend(counts[can_label] / counts[m] for m in alts)
        return np.array(weights, dtype=np.float32)

    def contains_modified_bases(self):
        return len(self.mod_long_names) > 0

    def __str__(self):
        s = 'canonical alphabet {}'.format(''.join(self.can_bases))
        if self.nmod_base == 0:
            return s + ' and no modified bases'
        return s + ' with modified base(s) ' + ', '.join(
            '{}={} (alt to {})'.format(m, self.mod_name_conv[m], c)
            for m, c in zip(self.alphabet, self.collapse_alphabet) if m in self.mod_bases_set)
