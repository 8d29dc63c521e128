"""Flip-flop coding utilities -- host-side mirror of taiyaki/flipflopfings.py
(same names and semantics).  The training ops build their transition indices on
the device (csrc/crf_flipflop.cu: indices_kernel); these numpy versions serve
batching code and tests."""
import numpy as np

DEFAULT_ALPHABET = 'ACGT'


def move_indices(labels, nbase=len(DEFAULT_ALPHABET)):
    """Transition index of each move labels[i] -> labels[i+1]
    (taiyaki/flipflopfings.py:6-17)."""
    nstate = nbase + nbase
    return labels[:-1] + np.minimum(labels[1:], nbase) * nstate


def stay_indices(labels, nbase=len(DEFAULT_ALPHABET)):
    """Transition index of staying in each label (flipflopfings.py:20-31)."""
    nstate = nbase + nbase
    return labels + np.minimum(labels, nbase) * nstate


def flopmask(labels):
    """True where a label sits at an even position of a homopolymer run
    (flipflopfings.py:34-53)."""
    move = np.ediff1d(labels, to_begin=1) != 0
    cumulative_flipflops = (1 - move).cumsum()
    offsets = np.maximum.accumulate(move * cumulative_flipflops)
    return (cumulative_flipflops - offsets) % 2 == 1


def flipflop_code(labels, alphabet_length=4):
    """Base labels -> flip-flop codes (flipflopfings.py:56-78)."""
    x = labels.copy()
    x[flopmask(x)] += alphabet_length
    return x


def flipflop_code_batch(labels, starts, alphabet_length=4):
    """`flipflop_code` of several concatenated label sequences in one pass;
    `starts` are the offsets of the sequences (a homopolymer run never crosses one)."""
    labels = np.asarray(labels)
    if labels.size == 0:
        return labels.astype(np.int64)
    move = np.ediff1d(labels, to_begin=1) != 0
    move[np.asarray(starts, dtype=np.int64)] = True
    cumulative_flipflops = (1 - move).cumsum()
    offsets = np.maximum.accumulate(move * cumulative_flipflops)
    flop = (cumulative_flipflops - offsets) % 2 == 1
    return (labels + alphabet_length * flop).astype(np.int64)


def nstate_flipflop(nbase):
    """Number of transitions 2L(L+1) (flipflopfings.py:146-168)."""
    return 2 * nbase * (nbase + 1)


def nbase_flipflop(nstate):
    """Inverse of nstate_flipflop, asserting validity (flipflopfings.py:171-184)."""
    nbase_f = np.sqrt(0.25 + (0.5 * np.float32(nstate))) - 0.5
    assert np.mod(nbase_f, 1) == 0, (
        'Number of states not valid for flip-flop model. ' +
        'nstates: {}\tconverted nbases: {}').format(nstate, nbase_f)
    return int(np.round(nbase_f))


def path_to_str(path, alphabet=DEFAULT_ALPHABET, include_first_source=True):
    """Flip-flop path -> basecall string (flipflopfings.py:81-99)."""
    move = np.ediff1d(path, to_begin=1 if include_first_source else 0) != 0
    alphabet = np.frombuffer((alphabet * 2).encode(), dtype='u1')
    seq = alphabet[path[move]]
    return seq.tobytes().decode()
