"""mLstm_cat_mod_flipflop: LSTM stack with the categorical modified-base
output layer (taiyaki models/mLstm_cat_mod_flipflop.py:6-19)."""
from taiyaki_b200.activation import swish
from taiyaki_b200.layers import Convolution, GlobalNormFlipFlopCatMod, Lstm, Reverse, Serial


def network(insize=1, size=256, winlen=19, stride=5, alphabet_info=None):
    winlen2 = 5
    return Serial([
        Convolution(insize, 4, winlen2, stride=1, fun=swish),
        Convolution(4, 16, winlen2, stride=1, fun=swish),
        Convolution(16, size, winlen, stride=stride, fun=swish),
        Reverse(Lstm(size, size)),
        Lstm(size, size),
        Reverse(Lstm(size, size)),
        Lstm(size, size),
        Reverse(Lstm(size, size)),
        GlobalNormFlipFlopCatMod(size, alphabet_info),
    ])
