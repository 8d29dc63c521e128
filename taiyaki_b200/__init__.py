"""taiyaki_b200 -- B200-native (sm_100a) flip-flop CRF training hot path.

Mirrors the operator surface of nanoporetech/taiyaki v5.3.0 for that path:

    taiyaki_b200.ctc.crf_flipflop_loss / cat_mod_flipflop_loss   (taiyaki/ctc/ctc.pyx)
    taiyaki_b200.layers.flipflop_logpartition, Lstm, GruMod, ... (taiyaki/layers.py)
    taiyaki_b200.flipflopfings                                   (taiyaki/flipflopfings.py)

All compute runs in hand-written CUDA kernels behind the C ABI declared in
include/taiyaki_b200.h; there is no CPU fallback.
"""
__version__ = '0.2.0'
