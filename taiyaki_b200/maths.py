"""Robust location / scale helpers of the training path, behind the reference's names
(taiyaki/maths.py:8-32 `med_mad`, :138-192 `RollingMAD`; same arguments and results).

Written for this package: `RollingMAD` keeps its history either in host memory (numpy, what
the reference does) or ON THE DEVICE (`device=`), where the thresholds of the gradient clipping
are then computed by two small sorts per step and never leave the GPU -- the train loop no
longer has to read the gradient maxima back before it can enqueue the next step
(bin/train_flipflop.py TrainLoop.run)."""
import numpy as np

MAD_SD_FACTOR = 1.4826          # MAD of a normal distribution -> its standard deviation


def _median(a, axis):
    return np.median(a, axis=axis, keepdims=True)


def med_mad(data, factor=None, axis=None, keepdims=False):
    """(median, factor * median absolute deviation) of `data` along `axis` (all elements when
    None).  `factor` defaults to the constant that makes the MAD estimate a normal's sd."""
    scale = MAD_SD_FACTOR if factor is None else factor
    data = np.asarray(data)
    centre = _median(data, axis)
    spread = scale * _median(np.abs(data - centre), axis)
    if axis is None:
        return centre.reshape(-1)[0], spread.reshape(-1)[0]
    if keepdims:
        return centre, spread
    return np.squeeze(centre, axis), np.squeeze(spread, axis)


def mad(data, factor=None, axis=None, keepdims=False):
    """The spread half of `med_mad` (taiyaki/maths.py:35-52)."""
    return med_mad(data, factor=factor, axis=axis, keepdims=keepdims)[1]


class RollingMAD:
    """Cap = median + n_mads * MAD over the last `window` updates, one cap per parameter.
    Until the window has been filled once, `update` returns `default_to`.

    device=None: numpy history, `update(values)` takes and returns host arrays.
    device=torch.device: history and caps are device tensors; `update(tensor)` returns a device
    tensor (or `default_to`) without synchronising."""

    def __init__(self, nparams, n_mads=0, window=1000, default_to=None, device=None):
        self.n_mads = n_mads
        self.default_to = default_to
        self.device = device
        self._seen = 0
        if device is None:
            self._history = np.empty((nparams, window), dtype='f4')
        else:
            import torch
            self._history = torch.empty((nparams, window), dtype=torch.float32, device=device)

    @property
    def nparams(self):
        return self._history.shape[0]

    @property
    def window(self):
        return self._history.shape[1]

    def update(self, vals):
        if len(vals) != self.nparams:
            raise AssertionError('Number of values ({}) provided does not match number of parameters ({}).'
                                 .format(len(vals), self.nparams))
        slot = self._seen % self.window
        self._seen += 1
        if self.device is None:
            self._history[:, slot] = vals
            if self._seen < self.window:
                return self.default_to
            centre, spread = med_mad(self._history, axis=1)
            return centre + self.n_mads * spread
        import torch
        self._history[:, slot] = vals.detach().to(self._history.dtype)
        if self._seen < self.window:
            return self.default_to
        # numpy's median (mean of the two middle values for an even window) as two sorts
        lo, hi = (self.window - 1) // 2, self.window // 2

        def median(t):
            s = torch.sort(t, dim=1).values
            return 0.5 * (s[:, lo] + s[:, hi])
        centre = median(self._history)
        spread = MAD_SD_FACTOR * median((self._history - centre[:, None]).abs())
        return centre + self.n_mads * spread
