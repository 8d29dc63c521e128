"""Building blocks shared by the model-definition files under models/: the
flip-flop networks of taiyaki (models/m{Lstm,Gru}[_cat_mod]_flipflop.py) are a
convolutional front-end, five recurrent layers of alternating direction
(backward first) and a flip-flop score layer."""
from . import layers
from .activation import swish, tanh


def alternating_stack(cell, size, depth=5):
    """`depth` recurrent layers of `cell`; even positions run backwards in time."""
    return [layers.Reverse(cell(size, size)) if i % 2 == 0 else cell(size, size)
            for i in range(depth)]


def lstm_front_end(insize, size, winlen, stride):
    """Three swish convolutions: 4 and 16 features at full rate, then the strided one."""
    small = 5
    return [layers.Convolution(insize, 4, small, stride=1, fun=swish),
            layers.Convolution(4, 16, small, stride=1, fun=swish),
            layers.Convolution(16, size, winlen, stride=stride, fun=swish)]


def gru_front_end(insize, size, winlen, stride):
    """One strided tanh convolution."""
    return [layers.Convolution(insize, size, winlen, stride=stride, fun=tanh)]


def score_layer(size, alphabet_info, cat_mod):
    if cat_mod:
        return layers.GlobalNormFlipFlopCatMod(size, alphabet_info)
    return layers.GlobalNormFlipFlop(size, 4 if alphabet_info is None else alphabet_info.nbase)
