"""Device-side batch assembly (csrc/batching.cu, taiyaki_b200/device_batching.py)
against the host path it replaces (chunk_selection.sample_chunks semantics,
signal_mapping.get_chunk_with_sample_length, Chunk.apply_filters, the stacking
and flip-flop coding of training.prepare_random_batches) on the SAME candidate
(read, start) list: identical chunk choice, labels and rejection counts, signal
to float32 round-off."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from taiyaki_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


def host_batch(reads, cands, N, T, fp, metadata, nbase):
    """What the reference's loop builds from this attempt sequence."""
    from collections import defaultdict
    from taiyaki_b200 import flipflopfings
    from taiyaki_b200.signal_mapping import Chunk
    chunks, rej = [], defaultdict(int)
    for r, s in zip(*cands):
        if len(chunks) >= N:
            break
        read = reads[r]
        if s < 0:
            chunk = Chunk(read.read_id, reject_reason=Chunk.rej_str_short)
        else:
            chunk = read.get_chunk_with_sample_length(
                T, start_sample=int(s) - read.get_mapped_dacs_region()[0],
                standardize=metadata.standardize)
        chunk.apply_filters(fp)
        rej[chunk.reject_reason] += 1
        if chunk.accepted:
            chunks.append(chunk)
    revop = np.flip if metadata.reverse else np.array
    cur = np.stack([revop(c.current) for c in chunks], 1).astype(np.float32)
    labels = [revop(c.sequence).astype(np.int64) for c in chunks]
    mods = None
    if metadata.is_cat_mod:
        mods = np.concatenate([metadata.mod_labels[x] for x in labels])
        labels = [metadata.can_labels[x] for x in labels]
    seqs = np.concatenate([flipflopfings.flipflop_code(np.ascontiguousarray(x), nbase)
                           for x in labels])
    return cur, seqs, np.array([len(x) for x in labels]), mods, dict(rej)


@pytest.mark.parametrize('T,N,reverse,standardize,cat_mod,use_filters', [
    (4000, 64, False, True, False, True),
    (1000, 16, True, True, False, True),
    (2000, 24, False, False, True, True),
    (500, 8, True, True, True, False),
    (50000, 4, False, True, False, True),        # reads shorter than the chunk: 'tooshort'
])
def test_device_batch_equals_host_batch(dev, T, N, reverse, standardize, cat_mod, use_filters):
    from taiyaki_b200 import chunk_selection, signal_mapping, training
    from taiyaki_b200.device_batching import DeviceReadStore
    np.random.seed(T + N)
    reads = signal_mapping.synthetic_reads(12, seed=3, mod_fraction=0.5 if cat_mod else 0.0)
    if cat_mod:
        md = training.NETWORK_METADATA(reverse, standardize, True, np.array([0, 1, 3, 4, 5]),
                                       np.array([0, 1, 2, 3, 1]), np.array([0, 0, 0, 0, 1]))
    else:
        md = training.NETWORK_METADATA(reverse, standardize, False)
    if use_filters and T <= 4000:
        fp = chunk_selection.sample_filter_parameters(reads, 50, T, 2.0, 3.0, 0.1, 5, 1.1)
    else:
        fp = chunk_selection.FILTER_PARAMETERS(10.0, 10.0, 0.1, None, None, None, None)
    store = DeviceReadStore(reads, dev)
    cands = store.draw_candidates(int(N / 0.1), T)
    if T > 40000:
        assert (cands[1] < 0).any()
    cur, seqs, seqlens, mods, rej = host_batch(reads, cands, N, T, fp, md, 4)
    indata, dseqs, dlens, dmods, n_acc, drej = store.sample(N, T, fp, md, 4, candidates=cands)
    torch.cuda.synchronize()
    assert n_acc == cur.shape[1] == len(seqlens)
    assert drej == rej
    assert indata.shape == (T, n_acc, 1)
    np.testing.assert_array_equal(dlens.cpu().numpy(), seqlens)
    np.testing.assert_array_equal(dseqs.cpu().numpy(), seqs)
    if cat_mod:
        np.testing.assert_array_equal(dmods.cpu().numpy(), mods)
    else:
        assert dmods is None
    # current: one fused multiply-add in fp32 against float64 arithmetic rounded to fp32
    np.testing.assert_allclose(indata[:, :, 0].cpu().numpy(), cur, rtol=2e-6, atol=2e-6)


def test_device_batches_train(dev):
    """The device generator feeds TrainStep like the host one."""
    import os
    from taiyaki_b200 import chunk_selection, device_batching, helpers, signal_mapping, training
    from taiyaki_b200.alphabet import AlphabetInfo
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    np.random.seed(0)
    torch.manual_seed(0)
    ai = AlphabetInfo('ACGT', 'ACGT')
    net = helpers.load_model(os.path.join(root, 'models', 'mLstm_flipflop.py'),
                             model_metadata={'reverse': False, 'standardize': True},
                             size=64, stride=5, winlen=19, insize=1, alphabet_info=ai).to(dev)
    net_info = training.NETWORK_INFO(net=net, net_clone=None,
                                     metadata=training.parse_network_metadata(net), stride=5)
    reads = signal_mapping.synthetic_reads(8, seed=2)
    fp = chunk_selection.sample_filter_parameters(reads, 50, 1000, 10.0, 10.0, 0.1, 5, 1.1)
    store = device_batching.DeviceReadStore(reads, dev)
    step = training.TrainStep(net_info, torch.optim.AdamW(net.parameters(), lr=2e-3, eps=1e-6))
    losses = []
    for _ in range(20):
        gen = device_batching.prepare_random_batches(store, 1000, 12, 1, ai, fp, net_info, None)
        _, loss, gmax = step(gen, sharpen=1.0)
        assert np.isfinite(loss)
        losses.append(loss)
    assert losses[-1] < losses[0]
