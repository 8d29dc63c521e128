"""mLstm_flipflop: 3 convolutions (swish) -> 5 alternating-direction LSTMs ->
flip-flop transition scores.  Same `network(...)` factory signature as
taiyaki's models/mLstm_flipflop.py:6-20."""
from taiyaki_b200 import layers, model_parts


def network(insize=1, size=256, winlen=19, stride=5, alphabet_info=None):
    return layers.Serial(model_parts.lstm_front_end(insize, size, winlen, stride) +
                         model_parts.alternating_stack(layers.Lstm, size) +
                         [model_parts.score_layer(size, alphabet_info, cat_mod=False)])
