"""A raw read as the data-preparation callers see it (taiyaki/signal.py:6-123): all DAC samples,
the channel's DAC -> pA conversion, and a trimmed view with the per-read shift / scale."""
import numpy as np

from . import fast5utils

NO_TRIM = {'trim_start': 0, 'trim_end': 0, 'shift': 0, 'scale': 1}
UNIT_CHANNEL = {'offset': 0, 'range': 1, 'digitisation': 1, 'sampling_rate': 4000}


class Signal:
    """Signal(read) takes samples, channel information and read id from a fast5 read
    (fast5utils.Fast5Read); Signal(dacs=array, channel_info=..., read_id=...) from memory.
    `read_params`: trim_start, trim_end, shift, scale of the read (one row of the per-read
    parameter table).  `untrimmed_dacs` keeps everything; `dacs`, `current` (pA) and
    `standardized_current` are the trimmed view."""

    def __init__(self, read=None, dacs=None, channel_info=UNIT_CHANNEL, read_id=None,
                 read_params=NO_TRIM):
        if read is None:
            if dacs is None:
                raise Exception('Cannot initialise Signal object')
            self.untrimmed_dacs = np.array(dacs)
            self.channel_info = channel_info
            self.read_id = read_id
        else:
            self.channel_info = dict(fast5utils.get_channel_info(read).items())
            read_id = fast5utils.get_read_attributes(read)['read_id']
            self.read_id = read_id.decode('utf-8') if isinstance(read_id, bytes) else str(read_id)
            self.untrimmed_dacs = read.get_raw_data()
        self.sample_rate = self.channel_info['sampling_rate']
        self.range = self.channel_info['range']
        self.offset = self.channel_info['offset']
        self.digitisation = self.channel_info['digitisation']
        self.set_trim_absolute(read_params['trim_start'], read_params['trim_end'])
        self.shift_from_pA = read_params['shift']
        self.scale_from_pA = read_params['scale']

    def set_trim_absolute(self, trimstart, trimend):
        """Trim counted from the ends of the whole read; a trim that would leave nothing is
        not applied (signal.py:77-95)."""
        if trimstart < 0 or trimend < 0:
            raise Exception("Can't trim a negative amount off the end of a signal vector.")
        n = len(self.untrimmed_dacs)
        if trimstart + trimend >= n:
            trimstart = trimend = 0
        self.signalstart, self.signalend_exc = trimstart, n - trimend

    def _pA(self, dacs):
        return (dacs + self.offset) * self.range / self.digitisation

    @property
    def dacs(self):
        return self.untrimmed_dacs[self.signalstart:self.signalend_exc].copy()

    @property
    def untrimmed_current(self):
        return self._pA(self.untrimmed_dacs)

    @property
    def current(self):
        return self._pA(self.dacs)

    @property
    def standardized_current(self):
        return (self.current - self.shift_from_pA) / self.scale_from_pA
