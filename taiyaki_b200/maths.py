"""Small numeric helpers on the training path (taiyaki/maths.py:8-32, :138-192)."""
import numpy as np

MAD_SD_FACTOR = 1.4826


def med_mad(data, factor=None, axis=None, keepdims=False):
    """Median and MAD (scaled to estimate a normal's sd), maths.py:8-32."""
    if factor is None:
        factor = MAD_SD_FACTOR
    dmed = np.median(data, axis=axis, keepdims=True)
    dmad = factor * np.median(abs(data - dmed), axis=axis, keepdims=True)
    if axis is None:
        dmed = dmed.flatten()[0]
        dmad = dmad.flatten()[0]
    elif not keepdims:
        dmed = dmed.squeeze(axis)
        dmad = dmad.squeeze(axis)
    return dmed, dmad


class RollingMAD:
    """Rolling median + n_mads * MAD cap over a window (maths.py:138-192)."""

    def __init__(self, nparams, n_mads=0, window=1000, default_to=None):
        self.n_mads = n_mads
        self.default_to = default_to
        self._window_data = np.empty((nparams, window), dtype='f4')
        self._curr_iter = 0

    @property
    def nparams(self):
        return self._window_data.shape[0]

    @property
    def window(self):
        return self._window_data.shape[1]

    def update(self, vals):
        assert len(vals) == self.nparams, (
            'Number of values ({}) provided does not match number of ' +
            'parameters ({}).').format(len(vals), self.nparams)
        self._window_data[:, self._curr_iter % self.window] = vals
        self._curr_iter += 1
        if self._curr_iter < self.window:
            return self.default_to
        med, mad = med_mad(self._window_data, axis=1)
        return med + (mad * self.n_mads)
