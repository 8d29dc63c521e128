"""mGru_cat_mod_flipflop: GRU stack with the categorical modified-base output
layer (taiyaki models/mGru_cat_mod_flipflop.py:6-15)."""
from taiyaki_b200.activation import tanh
from taiyaki_b200.layers import Convolution, GlobalNormFlipFlopCatMod, GruMod, Reverse, Serial


def network(insize=1, size=256, winlen=19, stride=2, alphabet_info=None):
    return Serial([
        Convolution(insize, size, winlen, stride=stride, fun=tanh),
        Reverse(GruMod(size, size)),
        GruMod(size, size),
        Reverse(GruMod(size, size)),
        GruMod(size, size),
        Reverse(GruMod(size, size)),
        GlobalNormFlipFlopCatMod(size, alphabet_info),
    ])
