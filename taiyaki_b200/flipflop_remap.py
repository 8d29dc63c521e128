"""Remapping of reads to their reference with a flip-flop model -- the interface of
taiyaki/flipflop_remap.py (map_to_crf_viterbi :6-86, flipflop_remap :89-143; SURVEY 8(f)
row 4) on the device kernel csrc/remap.cu.  `*_batch` variants align many reads in ONE
launch (one CTA per read), which is how prepare_mapped_reads-style jobs should call it;
the single-read functions keep the reference's signature and return types."""
import numpy as np
import torch

from . import _lib, flipflopfings

DEFAULT_ALPHABET = 'ACGT'
LARGE_VAL = 1e30


def remap_indices(sequence, alphabet=DEFAULT_ALPHABET):
    """(step_index, stay_index) of a base sequence (flipflop_remap.py:132-140)."""
    nbase = len(alphabet)
    lut = np.full(256, -1, dtype=np.int64)          # str.find: -1 for a letter not in the alphabet
    lut[np.frombuffer(alphabet.encode('latin-1'), dtype=np.uint8)[::-1]] = np.arange(nbase)[::-1]
    bases = lut[np.frombuffer(sequence.encode('latin-1'), dtype=np.uint8)]
    flops = flipflopfings.flopmask(bases)
    stay_index = np.where(flops, bases + (2 * nbase + 1) * nbase, bases + 2 * nbase * bases)
    from_base = (bases + flops * nbase)[:-1]
    to_base = np.maximum(bases, nbase * flops)[1:]
    step_index = from_base + 2 * nbase * to_base
    return step_index, stay_index


#: cap on the decision table (one byte per block x position) of one launch
MAX_TABLE_BYTES = 8 << 30


def launch_groups(table_bytes, cap=None):
    """Split reads (by their T x M decision-table sizes, in order) into consecutive groups
    whose tables fit `cap` bytes together; a read larger than the cap gets its own group."""
    cap = MAX_TABLE_BYTES if cap is None else cap
    groups, current, used = [], [], 0
    for r, nbytes in enumerate(table_bytes):
        if current and used + nbytes > cap:
            groups.append(current)
            current, used = [], 0
        current.append(r)
        used += nbytes
    if current:
        groups.append(current)
    return groups


def map_to_crf_viterbi_batch(scores_list, step_list, stay_list, localpen=LARGE_VAL):
    """Align several reads, one launch per group of reads whose decision tables fit
    MAX_TABLE_BYTES together (normally a single launch)."""
    sizes = [int(s.shape[0]) * len(st) for s, st in zip(scores_list, stay_list)]
    groups = launch_groups(sizes)
    if len(groups) <= 1:
        return _map_to_crf_viterbi_launch(scores_list, step_list, stay_list, localpen)
    out = []
    for grp in groups:
        out += _map_to_crf_viterbi_launch([scores_list[r] for r in grp], [step_list[r] for r in grp],
                                          [stay_list[r] for r in grp], localpen)
    return out


def _map_to_crf_viterbi_launch(scores_list, step_list, stay_list, localpen=LARGE_VAL):
    """Align several reads in one launch.  scores_list: [T_r, S] float arrays or tensors
    (host or device); step_list / stay_list: their index vectors.  Returns a list of
    (score, path[T_r + 1] int numpy array)."""
    nread = len(scores_list)
    assert nread > 0 and len(step_list) == nread and len(stay_list) == nread
    lib = _lib.lib()
    dev = None
    for s in scores_list:
        if isinstance(s, torch.Tensor) and s.is_cuda:
            dev = s.device
    if dev is None:
        if not torch.cuda.is_available():
            raise _lib.TaiyakiB200Error('flipflop_remap runs on the GPU only (no CPU fallback)')
        dev = torch.device('cuda', torch.cuda.current_device())
    S = int(scores_list[0].shape[1])
    T = np.array([int(s.shape[0]) for s in scores_list], dtype=np.int64)
    M = np.array([len(s) for s in stay_list], dtype=np.int64)
    for st, sp in zip(stay_list, step_list):
        assert len(sp) == len(st) - 1 and len(st) > 0
    assert all(int(s.shape[1]) == S for s in scores_list)
    t_off = np.concatenate([[0], np.cumsum(T)])
    m_off = np.concatenate([[0], np.cumsum(M)])
    tb_off = np.concatenate([[0], np.cumsum(T * M)])
    scores = torch.cat([torch.as_tensor(s).to(dev, torch.float32) for s in scores_list], 0).contiguous()
    stay = torch.as_tensor(np.concatenate([np.asarray(s) for s in stay_list]).astype(np.int32)).to(dev)
    step_np = np.concatenate([np.asarray(s, dtype=np.int64) for s in step_list]).astype(np.int32)
    step = torch.as_tensor(step_np if len(step_np) else np.zeros(1, dtype=np.int32)).to(dev)
    offs = torch.as_tensor(np.concatenate([t_off, m_off, tb_off])).to(dev)
    n1 = nread + 1
    score = torch.empty(nread, dtype=torch.float64, device=dev)
    path = torch.empty(int(t_off[-1]) + nread, dtype=torch.int32, device=dev)
    tb = torch.empty(max(int(tb_off[-1]), 1), dtype=torch.uint8, device=dev)
    max_m = int(M.max())
    dp = None
    if max_m > 12000:         # near / beyond the shared-memory capacity: global-memory score vectors
        dp = torch.empty(2 * int(m_off[-1]), dtype=torch.float64, device=dev)
    with _lib.timed('remap', dev):
        rc = lib.ty_flipflop_remap(_lib.ptr(scores), _lib.ptr(offs[:n1]), _lib.ptr(step),
                                   _lib.ptr(stay), _lib.ptr(offs[n1:2 * n1]), _lib.ptr(offs[2 * n1:]),
                                   nread, S, max_m, float(localpen), _lib.ptr(score), _lib.ptr(path),
                                   _lib.ptr(tb), _lib.ptr(dp), _lib.stream_ptr(dev))
    _lib.check(rc, 'ty_flipflop_remap')
    _lib.count_launches(1)
    score_h = score.cpu().numpy()
    path_h = path.cpu().numpy().astype(int)
    return [(score_h[r], path_h[t_off[r] + r:t_off[r + 1] + r + 1]) for r in range(nread)]


def map_to_crf_viterbi(scores, step_index, stay_index, localpen=LARGE_VAL):
    """Highest scoring path for a label sequence (flipflop_remap.py:6-86): returns
    (score of best path, best path)."""
    return map_to_crf_viterbi_batch([scores], [step_index], [stay_index], localpen)[0]


def flipflop_remap_batch(scores_list, sequences, alphabet=DEFAULT_ALPHABET, localpen=LARGE_VAL):
    idx = [remap_indices(seq, alphabet) for seq in sequences]
    return map_to_crf_viterbi_batch(scores_list, [i[0] for i in idx], [i[1] for i in idx], localpen)


def flipflop_remap(transition_scores, sequence, alphabet=DEFAULT_ALPHABET, localpen=LARGE_VAL):
    """Best alignment between flip-flop transition scores [T, S] and a sequence
    (flipflop_remap.py:89-143): (alignment score, positions array of length T + 1 with -1
    in the clipped start / end stretches)."""
    return flipflop_remap_batch([transition_scores], [sequence], alphabet, localpen)[0]
