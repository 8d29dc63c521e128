"""mGru_flipflop: convolution (tanh) -> 5 alternating-direction GRUs ->
flip-flop transition scores (taiyaki models/mGru_flipflop.py:6-17)."""
from taiyaki_b200 import layers, model_parts


def network(insize=1, size=256, winlen=19, stride=2, alphabet_info=None):
    return layers.Serial(model_parts.gru_front_end(insize, size, winlen, stride) +
                         model_parts.alternating_stack(layers.GruMod, size) +
                         [model_parts.score_layer(size, alphabet_info, cat_mod=False)])
