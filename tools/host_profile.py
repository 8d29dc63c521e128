#!/usr/bin/env python
"""cProfile of the host side of the train step at a launch-bound size (T_sig = 1000)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiyaki_b200 import chunk_selection, ctc, helpers, signal_mapping, training  # noqa: E402
from taiyaki_b200.alphabet import AlphabetInfo  # noqa: E402

dev = torch.device('cuda:0')
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ai = AlphabetInfo('ACGT', 'ACGT')
net = helpers.load_model(os.path.join(ROOT, 'models', 'mLstm_flipflop.py'),
                         model_metadata={'reverse': False, 'standardize': True}, stride=5, winlen=19,
                         insize=1, size=256, alphabet_info=ai).to(dev)
ni = training.NETWORK_INFO(net=net, net_clone=None, metadata=training.parse_network_metadata(net), stride=5)
opt = torch.optim.AdamW(net.parameters(), lr=1e-3, eps=1e-6, fused=True)
step = training.TrainStep(ni, opt)
reads = signal_mapping.synthetic_reads(24, seed=7)
fp = chunk_selection.sample_filter_parameters(reads, 100, T, 10.0, 10.0, 0.1, 5, 1.1)
b = list(training.prepare_random_batches(reads, T, 64, 1, ai, fp, ni, None))[0]
sl = b[2].to(dev)
ctc.hint_lengths(sl, int(b[2].max()), int(b[2].sum()))
batch = (b[0].to(dev), b[1].to(dev), sl, None, b[4], b[5])
for _ in range(5):
    step(iter([batch]), read_back=False)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step(iter([batch]), read_back=False)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
