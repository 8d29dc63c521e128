"""mGru_flipflop: convolution (tanh) -> 5 alternating-direction GRUs ->
flip-flop transition scores (taiyaki models/mGru_flipflop.py:6-17)."""
from taiyaki_b200.activation import tanh
from taiyaki_b200.layers import Convolution, GlobalNormFlipFlop, GruMod, Reverse, Serial


def network(insize=1, size=256, winlen=19, stride=2, alphabet_info=None):
    nbase = 4 if alphabet_info is None else alphabet_info.nbase
    return Serial([
        Convolution(insize, size, winlen, stride=stride, fun=tanh),
        Reverse(GruMod(size, size)),
        GruMod(size, size),
        Reverse(GruMod(size, size)),
        GruMod(size, size),
        Reverse(GruMod(size, size)),
        GlobalNormFlipFlop(size, nbase),
    ])
