#!/usr/bin/env python
"""Throughput of the basecalling driver (taiyaki_b200/basecall.py) on synthetic reads:
mLstm_flipflop 256, chunks of 1000 blocks with 100 blocks of overlap (the defaults of
bin/basecall.py), posterior-Viterbi decoding, fasta and fastq.  Host signals in, strings
out; one JSON line per configuration on stdout.

    python tools/basecall_bench.py [reads_per_batch ...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiyaki_b200 import (_lib, basecall, chunk_selection, device_batching, helpers,  # noqa: E402
                          signal_mapping, training)
from taiyaki_b200.alphabet import AlphabetInfo  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    _lib.lib()
    ai = AlphabetInfo('ACGT', 'ACGT')
    model = helpers.load_model(os.path.join(ROOT, 'models', 'mLstm_flipflop.py'),
                               model_metadata={'reverse': False, 'standardize': True}, stride=5,
                               winlen=19, insize=1, size=256, alphabet_info=ai).to(dev)
    stride = 5
    reads = signal_mapping.synthetic_reads(32, seed=11)
    # A random-weight network is nearly constant in time (the time variation of the
    # activations decays layer by layer), so its best path stays in one state and no bases
    # are called; the device work per block does not depend on that, the host's path_to_str
    # / quality-string work does.  TY_BENCH_TRAIN_STEPS > 0 trains on the same reads first
    # (thousands of steps are needed before the calls become realistic).
    np.random.seed(0)
    torch.manual_seed(0)
    net_info = training.NETWORK_INFO(net=model, net_clone=None,
                                     metadata=training.parse_network_metadata(model), stride=stride)
    fp = chunk_selection.sample_filter_parameters(reads, 100, 2000, 10.0, 10.0, 0.1, stride, 1.1)
    store = device_batching.DeviceReadStore(reads, dev)
    step = training.TrainStep(net_info, torch.optim.AdamW(model.parameters(), lr=2e-3, eps=1e-6))
    nstep = int(os.environ.get('TY_BENCH_TRAIN_STEPS', '0'))
    for it in range(nstep):
        gen = device_batching.prepare_random_batches(store, 2000, 48, 1, ai, fp, net_info, None)
        _, loss, _ = step(gen, sharpen=1.0)
    if nstep:
        print(json.dumps({'what': 'warm-up training', 'steps': nstep,
                          'final_loss': round(float(loss), 4)}), flush=True)
    signals = [(r.read_id, r.get_current(standardize=False).astype('f4')) for r in reads]
    nsample = sum(len(s) for _, s in signals)
    pools = [int(a) for a in sys.argv[1:]] or [1, 8, 32]
    for fastq in (False, True):
        for pool in pools:
            for concurrent in (128, 512):
                def run():
                    nbase = 0
                    for i in range(0, len(signals), pool):
                        for _, call, _, _ in basecall.process_signals(
                                signals[i:i + pool], model, 1000 * stride, 100 * stride, {}, 40, stride,
                                'ACGT', concurrent, fastq=fastq):
                            nbase += len(call)
                    return nbase
                run()
                torch.cuda.synchronize()
                t0 = time.time()
                nbase = run()
                torch.cuda.synchronize()
                dt = time.time() - t0
                print(json.dumps({'what': 'basecall', 'model': 'mLstm_flipflop 256', 'reads': len(signals),
                                  'samples': nsample, 'reads_per_batch': pool,
                                  'max_concurrent_chunks': concurrent, 'fastq': fastq,
                                  'seconds': round(dt, 4), 'Msamples_per_s': round(nsample / dt / 1e6, 3),
                                  'kbase_per_s': round(nbase / dt / 1e3, 1), 'bases_called': nbase}), flush=True)


if __name__ == '__main__':
    main()
