"""`taiyaki` import alias for the hot path rebuilt in `taiyaki_b200`.

The reference's model-definition files (`models/mLstm_flipflop.py:1-3`:
`from taiyaki.layers import ...`), its training script (`from taiyaki import
ctc, flipflopfings, ...`) and pickled `.checkpoint` files (class path
`taiyaki.layers.Serial`, taiyaki/helpers.py:104-113) name modules under
`taiyaki.`; with this package on the path they resolve to the B200
implementations without editing a line of them.  Only the modules on the
flip-flop training / basecalling path exist; anything else raises
ImportError naming what is missing rather than half-working.
"""
import importlib
import sys

import taiyaki_b200

__version__ = '5.3.0+b200.' + taiyaki_b200.__version__

#: taiyaki.<name> -> taiyaki_b200.<name>
ALIASED = ('activation', 'alphabet', 'basecall_helpers', 'chunk_selection', 'cmdargs', 'ctc', 'decode',
           'fast5utils', 'flipflop_remap', 'flipflopfings', 'helpers', 'json', 'layers', 'mapped_signal_files',
           'maths', 'prepare_mapping_funcs', 'qscores', 'signal', 'signal_mapping')

for _name in ALIASED:
    _mod = importlib.import_module('taiyaki_b200.' + _name)
    sys.modules[__name__ + '.' + _name] = _mod
    globals()[_name] = _mod
del _name, _mod


def __getattr__(name):
    raise ImportError(
        "taiyaki.{0} is outside the hot path rebuilt by taiyaki_b200 (available: {1})".format(
            name, ', '.join(ALIASED)))
