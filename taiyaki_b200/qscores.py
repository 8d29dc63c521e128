"""Quality scores of basecalls from posterior transition weights -- the interface
of taiyaki/qscores.py (qchar_from_qscore :10-28, qscore_from_errprob :31-40,
qchar_from_errprob :43-56, transitions_into_base :59-88, errprobs_from_trans
:91-140, path_errprobs_to_qstring :143-168).  The per-base loop of masked
matrix products becomes one [S] x [S, nbase] product with a constant matrix."""
import numpy as np
import torch

from . import flipflopfings

SMALL_VAL = 1e-10      # taiyaki/constants.py:7


def qchar_from_qscore(score, zerochar=33):
    """ASCII quality characters, score rounded to nearest (qscores.py:10-28)."""
    asciicodes = (np.array(score) + zerochar + 0.5).astype(np.int8)
    return asciicodes.tobytes().decode('ascii')


def qscore_from_errprob(errprob):
    """-10 log10(errprob) (qscores.py:31-40)."""
    return -10.0 * np.log10(errprob)


def qchar_from_errprob(errprob, qscore_scale, qscore_offset):
    """Quality characters of scale * q + offset (qscores.py:43-56)."""
    qscore = qscore_scale * qscore_from_errprob(errprob) + qscore_offset
    return qchar_from_qscore(qscore)


def transitions_into_base(b, nbases, device):
    """Indices of all transitions into base b, flip or flop (qscores.py:59-88)."""
    colstart = nbases * 2 * b
    toflip = torch.arange(colstart, colstart + nbases * 2, dtype=torch.long, device=device)
    fliptoflop = 2 * nbases * nbases + b
    toflop = torch.tensor([fliptoflop, fliptoflop + nbases], dtype=torch.long, device=device)
    return torch.cat((toflip, toflop))


def _into_base_matrix(nbases, device):
    """[S, nbases] 0/1 matrix: column b marks the transitions into base b."""
    nstate = flipflopfings.nstate_flipflop(nbases)
    m = np.zeros((nstate, nbases), dtype=np.float32)
    for b in range(nbases):
        m[nbases * 2 * b:nbases * 2 * (b + 1), b] = 1.0
        m[2 * nbases * nbases + b, b] = 1.0
        m[2 * nbases * nbases + nbases + b, b] = 1.0
    return torch.as_tensor(m, device=device)


def errprobs_from_trans(trans, path):
    """Error probability of every path element (qscores.py:91-140): one minus the
    posterior mass of the transitions into the called base over the mass of the
    transitions into any base.  trans [nblk, N, S] posterior weights (not logs),
    path [nblk+1, N]; returns [nblk+1, N] floats with -1 in row 0."""
    nblocks, batchsize, flip_flop_transitions = trans.shape
    nbases = flipflopfings.nbase_flipflop(flip_flop_transitions)
    baseprobs = torch.matmul(trans.float(), _into_base_matrix(nbases, trans.device))
    baseprobs = baseprobs / (baseprobs.sum(dim=2, keepdim=True) + SMALL_VAL)
    p = torch.empty_like(path, dtype=torch.float)
    ix = path[1:].unsqueeze(2) % nbases
    p[1:] = torch.gather(baseprobs, 2, ix).squeeze(2)
    p[0] = 2.0
    return 1.0 - p


def path_errprobs_to_qstring(errprobs, path, qscore_scale, qscore_offset):
    """Quality string of the emitted bases: stays and the source of the first
    transition are left out (qscores.py:143-168)."""
    filtered_probs = errprobs[1:][path[1:] != path[:-1]]
    if isinstance(filtered_probs, torch.Tensor):
        filtered_probs = filtered_probs.detach().cpu().numpy()
    return qchar_from_errprob(filtered_probs, qscore_scale, qscore_offset)
