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

    def contains_modified_bases(self):
        return len(self.mod_long_names) > 0

    def __str__(self):
        s = 'canonical alphabet {}'.format(''.join(self.can_bases))
        if self.nmod_base == 0:
            return s + ' and no modified bases'
        return s + ' with modified base(s) ' + ', '.join(
            '{}={} (alt to {})'.format(m, self.mod_name_conv[m], c)
            for m, c in zip(self.alphabet, self.collapse_alphabet) if m in self.mod_bases_set)
