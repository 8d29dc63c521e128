#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE itself.

Runs only in the build container (needs /root/reference):
  * imports the reference's Python (taiyaki.flipflopfings, taiyaki.layers,
    taiyaki.loss) from /root/reference for index coding, the TorchScript
    partition function + its autograd gradient, and FlipFlopLoss forward;
  * calls the reference's own C (oracle/_ref/libctc_ref.so, built by
    oracle/Makefile from /root/reference/taiyaki/ctc/*.c) for the
    label-constrained DP;
  * lifts the embedded known-answer tables out of the reference's C test mains
    (c_crf_flipflop.c:520-695, c_cat_mod_flipflop.c:586-870) as DATA.

The .npz files are committed; the GPU box never sees /root/reference.

    python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402
from oracle import oracle  # noqa: E402
from taiyaki import flipflopfings as ref_fff  # noqa: E402
from taiyaki import layers as ref_layers  # noqa: E402
from taiyaki import loss as ref_loss  # noqa: E402
from taiyaki.constants import SMALL_VAL  # noqa: E402


def c_array(src, name):
    """Pull `name[...] = { ... };` out of a C file as a flat float list."""
    m = re.search(r'\b' + name + r'\s*\[[^\]]*\]\s*=\s*\{(.*?)\};', src, re.S)
    assert m, name
    body = re.sub(r'//[^\n]*', '', m.group(1))
    return [float(x) for x in re.findall(r'[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?', body)]


def kat_tables():
    out = {}
    src = open(os.path.join(REF, 'taiyaki/ctc/c_crf_flipflop.c')).read()
    lp = np.log(np.array(c_array(src, 'test_logprob1'), dtype=np.float32)
                ).reshape(7, 2, 40).astype(np.float32)
    out['crf_logprob'] = lp
    out['crf_seq'] = np.array(c_array(src, 'test_seq1'), dtype=np.int64)
    out['crf_move'] = np.array(c_array(src, 'test_move1'), dtype=np.int64)
    out['crf_stay'] = np.array(c_array(src, 'test_stay1'), dtype=np.int64)
    out['crf_seqlen'] = np.array(c_array(src, 'test_seqlen1'), dtype=np.int64)
    # c_crf_flipflop.c:520-530 packs move with 5 entries per chunk
    mv = out['crf_move'][:10]
    st = out['crf_stay']
    sc, gr = oracle.c_crf_flipflop_grad(lp, mv, st, out['crf_seqlen'], 'ref')
    out['crf_move'] = mv
    out['crf_score'] = sc      # printed as -2.378088 -2.378088 by crf_test
    out['crf_grad'] = gr

    src = open(os.path.join(REF, 'taiyaki/ctc/c_cat_mod_flipflop.c')).read()
    main = src[src.index('#ifdef CAT_MOD_FLIPFLOP_TEST'):]
    names = re.findall(r'(\w+)\s*\[[^\]]*\]\s*=\s*\{', main)
    return out, main, names


def cat_mod_kat(main, names):
    """The cat-mod test main: tables + how main() calls the library."""
    out = {}
    for n in names:
        out['cm_' + n] = np.array(c_array(main, n))
    # main() (c_cat_mod_flipflop.c:799-823) takes log of the table and passes
    # the 12-entry index arrays as they are; the library offsets chunk 1's
    # move/mod arrays by seqidx - 1 = 5.
    lp = np.log(out['cm_test_logprob1'].astype(np.float32)).reshape(7, 2, 45)
    lp = lp.astype(np.float32)
    sl = out['cm_test_seqlen1'].astype(np.int64)
    mv = out['cm_test_move1'].astype(np.int64)
    st = out['cm_test_stay1'].astype(np.int64)
    mm = out['cm_test_modmoveidx1'].astype(np.int64)
    mf = out['cm_test_modmovefact1'].astype(np.float32)
    sc, gr = oracle.c_cat_mod_flipflop_grad(lp, mv, st, mm, mf, sl, 'ref')
    out = {'cm_logprob': lp, 'cm_seqlen': sl, 'cm_move': mv, 'cm_stay': st,
           'cm_modmove': mm, 'cm_modfact': mf, 'cm_score': sc, 'cm_grad': gr}
    # printed by cm_test as -52.354622 -195.435257
    return out


def ctc_loss_unit_case():
    """test/unit/test_ctc_loss.py:24-103 -- scores whose path probabilities
    are known: P(015) = P(237) = 0.5, P(510) ~ 0; logZ = 0."""
    nbases, nblocks = 4, 4

    def tc(f, t):
        return t * 2 * nbases + f if t < nbases else 2 * nbases * nbases + f
    paths = {'015': [0, 0, 1, 5, 5], '237': [2, 2, 3, 7, 7]}
    weights = {'015': [1.0, 1.0, 0.5, 1.0], '237': [1.0, 0.5, 1.0, 1.0]}
    outputs = torch.zeros(nblocks, 1, 40, dtype=torch.float)
    for k in paths:
        for blk in range(nblocks):
            outputs[blk, 0, tc(paths[k][blk], paths[k][blk + 1])] = weights[k][blk]
    outputs = torch.log(outputs + SMALL_VAL)
    outputs = ref_layers.global_norm_flipflop(outputs)
    logpart = float(ref_layers.log_partition_flipflop(outputs))
    return {'unit_scores': outputs.numpy(),
            'unit_logpart': np.float32(logpart),
            'unit_seqs': np.array([[0, 1, 5], [2, 3, 7], [5, 1, 0]], dtype=np.int64),
            'unit_probs': np.array([0.5, 0.5, 0.0], dtype=np.float64)}


def ref_indices(seqs, seqlen, nbase):
    parts = np.split(seqs.astype(np.int32), np.cumsum(seqlen[:-1]))
    mv = np.concatenate([ref_fff.move_indices(s, nbase) for s in parts])
    st = np.concatenate([ref_fff.stay_indices(s, nbase) for s in parts])
    return mv.astype(np.int64), st.astype(np.int64)


def random_case(tag, nblk, lengths, ntrans, seed, sharp=1.0, scale=1.0):
    nbatch = len(lengths)
    out = {}
    scores = oracle.synth_scores(nblk, nbatch, ntrans, seed=seed) * np.float32(scale)
    rng = np.random.RandomState(seed + 100)
    raw = [rng.randint(0, 4, size=int(L)) for L in lengths]
    seqs = np.concatenate([ref_fff.flipflop_code(r, 4) if len(r) else r.astype(np.int64)
                           for r in raw]).astype(np.int64)
    seqlen = np.array(lengths, dtype=np.int64)
    mv, st = ref_indices(seqs, seqlen, 4)
    out['scores'] = scores
    out['seqs'] = seqs
    out['seqlen'] = seqlen
    out['move'] = mv
    out['stay'] = st
    out['sharp'] = np.float32(sharp)
    if ntrans == 40:
        lp = np.float32(sharp) * scores
        sc, gr = oracle.c_crf_flipflop_grad(lp, mv, st, seqlen, 'ref')
        out['score'] = sc
        out['grad'] = gr
        out['score_costonly'] = oracle.c_crf_flipflop_cost(lp, mv, st, seqlen, 'ref')
        # secondary forward-only oracle: taiyaki/loss.py:113-173
        if all(L > 0 for L in lengths):
            try:
                ffl = ref_loss.FlipFlopLoss(sharp=float(sharp))
                Lmax = int(max(lengths))
                pst = np.zeros((nbatch, Lmax), dtype=np.int64)
                pmv = np.zeros((nbatch, Lmax - 1), dtype=np.int64)
                for b, s in enumerate(np.split(seqs, np.cumsum(seqlen[:-1]))):
                    pst[b, :len(s)] = ref_fff.stay_indices(s, 4)
                    pmv[b, :len(s) - 1] = ref_fff.move_indices(s, 4)
                with torch.no_grad():
                    v = ffl(torch.tensor(scores), torch.tensor(pmv),
                            torch.tensor(pst), torch.tensor(seqlen))
                out['torch_flipfloploss'] = np.asarray(v, dtype=np.float32)
            except Exception as e:  # informational only
                out['torch_flipfloploss_err'] = np.array(str(e))
    else:
        can_mods_offsets = np.array([0, 1, 3, 4, 5], dtype=np.int32)
        mod_cats = np.concatenate([
            ((r == 1) & (rng.uniform(size=len(r)) < 0.5)).astype(np.int64)
            for r in raw]).astype(np.int64)
        w = np.array([1.0, 1.0, 0.7, 1.0, 1.0], dtype=np.float32)
        trans_sharp = np.ones(ntrans, dtype=np.float32)
        trans_sharp[:40] = sharp
        lp = np.ascontiguousarray(scores * trans_sharp)
        mm, mf = oracle.build_mod_indices(seqs, seqlen, mod_cats, can_mods_offsets, w, 4)
        sc, gr = oracle.c_cat_mod_flipflop_grad(lp, mv, st, mm, mf, seqlen, 'ref')
        out['mod_cats'] = mod_cats
        out['can_mods_offsets'] = can_mods_offsets
        out['mod_cat_weights'] = w
        out['modmove'] = mm.astype(np.int64)
        out['modfact'] = mf
        out['score'] = sc
        out['grad'] = gr
        out['score_costonly'] = oracle.c_cat_mod_flipflop_cost(
            lp, mv, st, mm, mf, seqlen, 'ref')
    # partition function + autograd gradient from the reference's TorchScript
    x = torch.tensor(scores[:, :, :40].copy(), requires_grad=True)
    lz = ref_layers.log_partition_flipflop(x).squeeze(1)
    lz.sum().backward()
    out['logz'] = lz.detach().numpy()
    out['logz_grad'] = x.grad.numpy()
    return {tag + '_' + k: v for k, v in out.items()}


def decodeutil_case():
    """test/unit/test_decodeutil.py:16-31: seed 0xdeadbeef, randn(12,40)."""
    np.random.seed(0xdeadbeef)
    w = np.random.randn(12, 40).astype('f4')
    lz = float(ref_layers.log_partition_flipflop(torch.tensor(w).unsqueeze(1)))
    return {'du_weights': w, 'du_logz_flipstart': np.float32(lz),
            'du_free_start_expected': np.float64(27.16876983642578)}


def flipflop_code_cases():
    rng = np.random.RandomState(5)
    out = {}
    ex = np.array([1, 3, 2, 3, 3, 3, 3, 1, 1])
    out['code_in0'] = ex
    out['code_out0'] = ref_fff.flipflop_code(ex)
    x = rng.randint(0, 4, size=400)
    x[50:60] = 2
    out['code_in1'] = x
    out['code_out1'] = ref_fff.flipflop_code(x)
    out['code_move1'] = ref_fff.move_indices(out['code_out1'], 4)
    out['code_stay1'] = ref_fff.stay_indices(out['code_out1'], 4)
    return out


def main():
    oracle.build()
    assert oracle.have_ref()
    g = {}
    kat, cm_main, cm_names = kat_tables()
    g.update(kat)
    g.update(cat_mod_kat(cm_main, cm_names))
    g.update(ctc_loss_unit_case())
    g.update(decodeutil_case())
    g.update(flipflop_code_cases())
    np.savez_compressed(os.path.join(HERE, 'kat.npz'), **g)

    r = {}
    r.update(random_case('a', 50, [20, 1, 27, 13, 0, 45, 2], 40, seed=11))
    r.update(random_case('b', 64, [30, 35, 28], 40, seed=12, sharp=2.0))
    r.update(random_case('c', 40, [18, 22, 1, 0, 9], 45, seed=13))
    r.update(random_case('d', 33, [15, 17], 45, seed=14, sharp=1.5))
    r.update(random_case('e', 120, [100, 64, 33, 120], 40, seed=15, scale=4.0))
    np.savez_compressed(os.path.join(HERE, 'random.npz'), **r)
    for f in ('kat.npz', 'random.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')
    print('crf KAT score', g['crf_score'])
    print('cat-mod tables:', cm_names)


def make_decode():
    """Viterbi paths and posterior transition probabilities of the reference's own
    PyTorch implementations (taiyaki/decode.py:_flipflop_viterbi, flipflop_make_trans
    with _never_use_cupy=True) on seeded random scores -> tests/golden/decode.npz."""
    from taiyaki import decode as ref_decode
    out = {}
    for tag, (T, N, seed, scale) in {'a': (50, 3, 11, 1.0), 'b': (200, 3, 12, 5.0),
                                     'c': (1, 2, 13, 1.0)}.items():
        g = torch.Generator().manual_seed(seed)
        scores = scale * torch.randn(T, N, 40, generator=g)
        fwd, tb, path = ref_decode._flipflop_viterbi(scores)
        trans = ref_decode.flipflop_make_trans(scores, _never_use_cupy=True)
        out[tag + '_scores'] = scores.numpy()
        out[tag + '_fwd'] = fwd.numpy()
        out[tag + '_tb'] = tb.numpy().astype(np.int8)
        out[tag + '_path'] = path.numpy().astype(np.int8)
        out[tag + '_trans'] = trans.numpy()
    np.savez_compressed(os.path.join(HERE, 'decode.npz'), **out)
    print('decode.npz', {k: v.shape for k, v in out.items() if k.startswith('b_')})


def make_basecall():
    """Chunking / stitching / quality strings / path -> bases of the reference's own
    Python (taiyaki/basecall_helpers.py, qscores.py, flipflopfings.path_to_str) and the
    post-network flow of bin/basecall.py:216-243 (process_read after the model call)
    on seeded inputs -> tests/golden/basecall.npz.

    Two shims, neither touching the arithmetic: `taiyaki.helpers` cannot be imported
    under Python 3.12 (`import imp`), so a stub module with the two names
    basecall_helpers imports is registered first; `qscores.qchar_from_qscore` calls
    ndarray.tostring(), removed in numpy 2 -- it is replaced by the same expression
    with tobytes()."""
    import types
    stub = types.ModuleType('taiyaki.helpers')
    stub.get_model_device = stub.guess_model_stride = None
    sys.modules.setdefault('taiyaki.helpers', stub)
    from taiyaki import basecall_helpers as ref_bh
    from taiyaki import decode as ref_decode
    from taiyaki import qscores as ref_q

    def qchar_from_qscore(score, zerochar=33):
        asciicodes = (np.array(score) + zerochar + 0.5).astype(np.int8)
        return asciicodes.tobytes().decode('ascii')
    ref_q.qchar_from_qscore = qchar_from_qscore

    out = {}
    rng = np.random.RandomState(21)
    cases = {'a': (10000, 1000, 100, 5), 'b': (1000, 1000, 100, 5), 'c': (999, 1000, 100, 5),
             'd': (2345, 500, 0, 2), 'e': (5003, 1000, 500, 5), 'f': (1001, 1000, 100, 2),
             'g': (7777, 800, 700, 4)}
    for tag, (nsample, chunk_size, overlap, stride) in cases.items():
        signal = rng.standard_normal(nsample).astype('f4')
        chunks, cs, ce = ref_bh.chunk_read(signal, chunk_size, overlap)
        out[tag + '_cfg'] = np.array([nsample, chunk_size, overlap, stride])
        out[tag + '_signal'] = signal
        out[tag + '_chunks'] = chunks
        out[tag + '_starts'] = cs
        out[tag + '_ends'] = ce
        T = chunks.shape[0] // stride
        net_out = torch.tensor(rng.standard_normal((T, chunks.shape[1], 3)).astype('f4'))
        path = torch.tensor(rng.randint(0, 8, size=(T + 1, chunks.shape[1])))
        out[tag + '_out'] = net_out.numpy()
        out[tag + '_path'] = path.numpy().astype(np.int8)
        out[tag + '_stitched'] = ref_bh.stitch_chunks(net_out, cs, ce, stride).numpy()
        out[tag + '_stitched_path'] = ref_bh.stitch_chunks(path, cs, ce, stride).numpy().astype(np.int8)
        out[tag + '_stitched_path_ps'] = ref_bh.stitch_chunks(
            path, cs, ce, stride, path_stitching=True).numpy().astype(np.int8)

    # post-network flow of process_read: scores of the chunks of a 9000-sample read
    stride, chunk_size, overlap, nsample = 5, 1000, 100, 9000
    _, cs, ce = ref_bh.chunk_read(np.zeros(nsample, dtype='f4'), chunk_size, overlap)
    g = torch.Generator().manual_seed(31)
    scores = 3.0 * torch.randn(chunk_size // stride, len(cs), 40, generator=g)
    out['flow_scores'] = scores.numpy()
    out['flow_starts'], out['flow_ends'] = cs, ce
    out['flow_cfg'] = np.array([nsample, chunk_size, overlap, stride])
    for tag, posterior, temperature in (('post', True, 1.0), ('raw', False, 1.0), ('temp', True, 0.5)):
        trans = scores * temperature
        if posterior:
            trans = (ref_decode.flipflop_make_trans(trans, _never_use_cupy=True) + 1e-8).log()
        _, _, chunk_best_paths = ref_decode._flipflop_viterbi(trans)
        best_path = ref_bh.stitch_chunks(chunk_best_paths, cs, ce, stride).numpy()
        basecall = ref_fff.path_to_str(best_path, alphabet='ACGT', include_first_source=False)
        if posterior:
            # process_read hands the LOG posterior weights to errprobs_from_trans
            # (bin/basecall.py:216-217, 231-232); without --posterior the raw scores give
            # negative "probabilities" and NaN quality values, so no vector is kept for that
            chunk_errprobs = ref_q.errprobs_from_trans(trans, chunk_best_paths)
            errprobs = ref_bh.stitch_chunks(chunk_errprobs, cs, ce, stride)
            assert bool(((errprobs[1:] > 0) & (errprobs[1:] < 1)).all())
            qstring = ref_q.path_errprobs_to_qstring(errprobs, best_path, 1.0, 0.0)
            assert len(basecall) == len(qstring)
            out['flow_%s_errprobs' % tag] = errprobs.numpy()
            out['flow_%s_qstring' % tag] = np.array(qstring)
        if tag == 'post':
            out['flow_post_trans'] = trans.numpy()
        out['flow_%s_chunk_paths' % tag] = chunk_best_paths.numpy().astype(np.int8)
        out['flow_%s_path' % tag] = best_path.astype(np.int8)
        out['flow_%s_basecall' % tag] = np.array(basecall)
    # errprobs_from_trans on genuine posterior weights (the documented use)
    post = ref_decode.flipflop_make_trans(scores[:, :3], _never_use_cupy=True)
    _, _, paths = ref_decode._flipflop_viterbi(scores[:, :3])
    out['q_trans'] = post.numpy()
    out['q_paths'] = paths.numpy().astype(np.int8)
    out['q_errprobs'] = ref_q.errprobs_from_trans(post, paths).numpy()
    out['q_qchars'] = np.array(ref_q.qchar_from_errprob(np.array([0.5, 0.1, 0.011, 1e-4, 0.9999]), 1.0, 0.0))
    out['q_qchars_cal'] = np.array(ref_q.qchar_from_errprob(np.array([0.5, 0.1, 0.011, 1e-4]), 0.9, 1.5))
    np.savez_compressed(os.path.join(HERE, 'basecall.npz'), **out)
    print('basecall.npz', os.path.getsize(os.path.join(HERE, 'basecall.npz')), 'bytes;',
          'flow basecall lengths', [len(str(out['flow_%s_basecall' % t])) for t in ('post', 'raw', 'temp')])


def make_real():
    """Real r9.4.1 reads (the reference's test/data/mapped_signal_file/mapped_reads_{0,1}.hdf5,
    7 reads from the walkthrough data set) -> tests/golden/real_reads.npz:

      * the arrays and attributes of every read, as decoded by taiyaki_b200/hdf5_min.py (the
        reference cannot read the files here either -- no h5py); each decoded read is built
        into the REFERENCE's signal_mapping.SignalMapping and must pass its own check();
      * the reference's chunk sampling on them (chunk_selection.sample_filter_parameters and
        sample_chunks under a fixed numpy seed): filter parameters, and per chunk the read,
        start sample, standardised current, sequence, dwell statistics, rejection reason;
      * the reference's C loss (crf_flipflop_grad via oracle/_ref) on seeded random scores
        with the flip-flop coded REAL label sequences of the accepted chunks."""
    from taiyaki import chunk_selection as ref_cs
    from taiyaki import signal_mapping as ref_sm
    from taiyaki_b200 import hdf5_min
    out = {}
    reads = []
    names = []
    for fn in ('mapped_reads_0', 'mapped_reads_1'):
        f = hdf5_min.File(os.path.join(REF, 'test/data/mapped_signal_file', fn + '.hdf5'))
        assert int(f.attrs['version']) == 8 and f.attrs['alphabet'] == 'ACGT'
        out[fn + '_read_ids'] = np.array(f['Reads'].keys())
        for rid in f['Reads'].keys():
            g = f['Reads/' + rid]
            d = {k: g[k][()] for k in g.keys()}
            d.update(g.attrs)
            assert d['read_id'] == rid
            sm = ref_sm.SignalMapping(**d)
            assert sm.check() == ref_sm.SignalMapping.pass_str, sm.check()
            reads.append(sm)
            names.append(rid)
            for k in ('Dacs', 'Ref_to_signal', 'Reference'):
                out['read_%s_%s' % (rid, k)] = d[k]
            out['read_%s_attrs' % rid] = np.array([d['shift_frompA'], d['scale_frompA'], d['range'],
                                                    d['offset'], d['digitisation']], dtype=np.float64)
    out['read_ids'] = np.array(names)
    chunk_len, stride, nsample = 1000, 5, 40
    np.random.seed(17)
    fp = ref_cs.sample_filter_parameters(reads, 100, chunk_len, 10.0, 10.0, 0.1, stride, 1.1)
    out['fp_median_meandwell'] = np.float64(fp.median_meandwell)
    out['fp_mad_meandwell'] = np.float64(fp.mad_meandwell)
    out['fp_args'] = np.array([100, chunk_len, 10.0, 10.0, 0.1, stride, 1.1])
    np.random.seed(18)
    # tighter filters than the defaults so that some real chunks are rejected
    fp2 = fp._replace(filter_mean_dwell=1.5, filter_max_dwell=6.0)
    chunks, rejections = ref_cs.sample_chunks(reads, nsample, chunk_len, fp2)
    out['chunk_seed'] = np.array([17, 18])
    out['chunk_filter'] = np.array([1.5, 6.0])
    out['chunk_read'] = np.array([names.index(c.read_id) for c in chunks])
    out['chunk_start'] = np.array([c.start_sample for c in chunks])
    out['chunk_current'] = np.stack([c.current for c in chunks]).astype(np.float64)
    out['chunk_seqlen'] = np.array([len(c.sequence) for c in chunks])
    out['chunk_seq'] = np.concatenate([c.sequence for c in chunks]).astype(np.int16)
    out['chunk_mean_dwell'] = np.array([c.mean_dwell for c in chunks], dtype=np.float64)
    out['chunk_max_dwell'] = np.array([c.max_dwell for c in chunks], dtype=np.float64)
    out['rejection_keys'] = np.array(sorted(rejections))
    out['rejection_counts'] = np.array([rejections[k] for k in sorted(rejections)])
    # loss on the real label sequences of the first 8 accepted chunks
    nb = 8
    nblk = chunk_len // stride
    seqs = np.concatenate([ref_fff.flipflop_code(c.sequence.astype(np.int64), 4)
                           for c in chunks[:nb]]).astype(np.int64)
    seqlen = out['chunk_seqlen'][:nb].astype(np.int64)
    scores = oracle.synth_scores(nblk, nb, 40, seed=41)
    mv, st = ref_indices(seqs, seqlen, 4)
    sc, gr = oracle.c_crf_flipflop_grad(scores, mv, st, seqlen, 'ref')
    out['loss_scores'] = scores
    out['loss_seqs'] = seqs
    out['loss_seqlen'] = seqlen
    out['loss_score'] = sc
    out['loss_grad'] = gr
    np.savez_compressed(os.path.join(HERE, 'real_reads.npz'), **out)
    print('real_reads.npz', os.path.getsize(os.path.join(HERE, 'real_reads.npz')), 'bytes;',
          len(reads), 'reads;', len(chunks), 'chunks; rejections', dict(rejections),
          '; median/mad mean dwell', fp.median_meandwell, fp.mad_meandwell, '; L', seqlen)


def make_trained():
    """The reference's SHIPPED trained model on REAL signal -> tests/golden/trained_r941.npz.

    models/mLstm_flipflop_model_r941_DNA.checkpoint (mLstm_flipflop, size 256, stride 5: the
    architecture of BASELINE configs[1]) is unpickled into the REFERENCE's own taiyaki.layers
    classes (stock nn.LSTM / nn.Conv1d / nn.Linear; the one shim is the `_flat_weights` list
    torch >= 1.8 expects and a torch-1.5 pickle lacks) and run on the CPU in fp32 over the real
    r9.4.1 chunks of real_reads.npz.  Every parameter is rounded to bf16 first -- on BOTH sides:
    the rounded values are what is stored (uint16, half the bytes) and what the reference
    network computes with -- so the comparison isolates the arithmetic of the layers from the
    storage format of the fixture.  Stored: the parameters, the reference's scores
    [200, 16, 40], and the reference loss (its C via oracle/_ref + its TorchScript partition
    function) of each chunk against its real label sequence."""
    import warnings
    from taiyaki_b200.helpers import _legacy_rnn_pickles
    real = np.load(os.path.join(HERE, 'real_reads.npz'))
    with _legacy_rnn_pickles(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net = torch.load(os.path.join(REF, 'models/mLstm_flipflop_model_r941_DNA.checkpoint'),
                         map_location='cpu', weights_only=False)
    assert type(net).__module__ == 'taiyaki.layers' and ref_layers.Serial is type(net)
    out = {}
    with torch.no_grad():
        for name, p in net.state_dict().items():
            r = p.detach().to(torch.bfloat16)
            p.copy_(r.float())
            out['param_' + name] = r.view(torch.int16).numpy().view(np.uint16)
    nb = 16
    x = torch.tensor(real['chunk_current'][:nb].T.astype(np.float32)).unsqueeze(2)   # [1000, 16, 1]
    net.eval()
    with torch.no_grad():
        scores = net(x)
    assert scores.shape == (200, nb, 40)
    out['signal'] = x.numpy()
    out['scores'] = scores.numpy()
    seqlen = real['chunk_seqlen'][:nb].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(real['chunk_seqlen'])])
    seqs = np.concatenate([ref_fff.flipflop_code(real['chunk_seq'][off[i]:off[i + 1]].astype(np.int64), 4)
                           for i in range(nb)]).astype(np.int64)
    mv, st = ref_indices(seqs, seqlen, 4)
    sc, gr = oracle.c_crf_flipflop_grad(scores.numpy(), mv, st, seqlen, 'ref')
    logz = ref_layers.log_partition_flipflop(scores).squeeze(1).numpy()
    nblk = scores.shape[0]
    out['seqs'], out['seqlen'] = seqs, seqlen
    out['loss'] = (-sc / nblk + logz / nblk).astype(np.float32)     # train_flipflop.py:163-176
    np.savez_compressed(os.path.join(HERE, 'trained_r941.npz'), **out)
    print('trained_r941.npz', os.path.getsize(os.path.join(HERE, 'trained_r941.npz')), 'bytes; loss per chunk',
          np.round(out['loss'], 4), '; score range', float(scores.min()), float(scores.max()))


def make_remap():
    """Alignments of the reference's taiyaki/flipflop_remap.py on seeded random scores and
    on the two tables of its unit test (test/unit/test_flipflop_remap.py:8-90)
    -> tests/golden/remap.npz.

    Shim: under numpy 2 the traceback line `m -= move` (flipflop_remap.py:85) wraps the
    uint8 `move` around instead of reaching -1 when the path leaves through the start
    state, and the function raises IndexError.  The module source is loaded with that one
    statement changed to `m -= int(move)`; nothing else differs from the file on disk.

    Precision: the reference pins numpy 1.18 (requirements.txt:9), where the scalar updates
    `start_score + max(stay_scores[0], -localpen)` (:57, :67) promote the fp32 score to
    float64 like the vector updates do.  numpy 2 keeps such scalar sums in float32.  The fp32
    scores are therefore handed over as float64 (an exact conversion), which gives the
    pinned-numpy arithmetic -- all float64 -- under either numpy."""
    import types
    src = open(os.path.join(REF, 'taiyaki/flipflop_remap.py')).read()
    assert src.count('m -= move') == 1
    ref = types.ModuleType('ref_flipflop_remap')
    exec(compile(src.replace('m -= move', 'm -= int(move)'), 'flipflop_remap.py', 'exec'), ref.__dict__)
    out = {}
    rng = np.random.RandomState(51)
    cases = {'a': (6, 4, -0.5, 3.0), 'b': (50, 12, 1e30, 3.0), 'c': (80, 10, 0.5, 3.0),
             'd': (200, 60, 2.0, 3.0), 'e': (40, 1, 0.3, 3.0), 'f': (30, 29, 1e30, 3.0),
             'g': (300, 40, 0.0, 3.0), 'h': (300, 100, 1.0, 5.0), 'i': (120, 30, 0.2, 1.0),
             'j': (257, 129, 1e30, 2.0), 'k': (64, 33, 0.7, 0.5)}
    for tag, (T, L, pen, scale) in cases.items():
        seq = ''.join('ACGT'[i] for i in rng.randint(0, 4, size=L))
        if tag in 'hj':        # homopolymer runs exercise the flop coding
            seq = ''.join(c * int(k) for c, k in zip(seq[:L // 2], rng.randint(1, 4, size=L // 2)))[:L]
        sc = (scale * rng.standard_normal((T, 40))).astype('f4')
        score, path = ref.flipflop_remap(sc.astype('f8'), seq, localpen=pen)
        out[tag + '_scores'] = sc
        out[tag + '_seq'] = np.array(seq)
        out[tag + '_localpen'] = np.float64(pen)
        out[tag + '_score'] = np.float64(score)
        out[tag + '_path'] = path.astype(np.int32)
        # block path -> Ref_to_signal: the arithmetic of SignalMapping.from_remapping_path
        # (signal_mapping.py:303-318) with the reference's own get_reftosignal; stride 5, a
        # signal trimmed by 7 samples at the start, 13 spare samples at the end
        from taiyaki.signal_mapping import SignalMapping as RefSM
        stride, signalstart = 5, 7
        nd = T * stride + signalstart + 13
        full = np.full(nd, -1, dtype=np.int32)
        siglocs = np.arange(len(path), dtype=np.int32) * stride - 1 + signalstart
        f = np.logical_and(siglocs >= 0, siglocs < nd)
        full[siglocs[f]] = path[f]
        out[tag + '_reftosig'] = RefSM.get_reftosignal(full, len(seq), nd)
        out[tag + '_reftosig_cfg'] = np.array([stride, signalstart, nd])
    # unit-test tables: alphabet AB, 12 transitions; expected values are the test's own
    lt = np.zeros((6, 12), dtype='f4')
    for t, k in enumerate((8, 10, 6, 5, 1, 0)):
        lt[t, k] = 1
    score, path = ref.flipflop_remap(lt, 'AABA', alphabet='AB', localpen=-0.5)
    assert score == 6.0 and path.tolist() == [0, 1, 1, 2, 2, 3, 3]
    out['kat1_scores'], out['kat1_seq'], out['kat1_score'], out['kat1_path'] = lt, np.array('AABA'), score, path
    lt = np.zeros((5, 12), dtype='f4')
    lt[2, 5] = 1
    lt[3, 1] = 1
    score, path = ref.flipflop_remap(lt, 'BA', alphabet='AB', localpen=-0.5)
    assert score == 3.5 and path.tolist() == [-1, -1, 0, 0, 1, -1]
    out['kat2_scores'], out['kat2_seq'], out['kat2_score'], out['kat2_path'] = lt, np.array('BA'), score, path
    np.savez_compressed(os.path.join(HERE, 'remap.npz'), **out)
    print('remap.npz', os.path.getsize(os.path.join(HERE, 'remap.npz')), 'bytes;',
          {t: (float(out[t + '_score']), int((out[t + '_path'] == -1).sum())) for t in cases})


def make_prepare():
    """The reference's data-preparation flow on its own raw-read fixtures -> tests/golden/prepare_remap.npz.

    Inputs (all from /root/reference): the five single-read fast5 files of test/data/reads
    (decoded with taiyaki_b200/hdf5_min.py -- no h5py / ont_fast5_api here; pinned below against
    what the reference itself stored from the same files), test/data/readparams.tsv,
    test/data/per_read_references.fasta and the shipped remapping model
    models/mGru_flipflop_remapping_model_r9_DNA.checkpoint (mGru_flipflop, size 96, stride 4).

    Computed with the REFERENCE's code on the CPU in fp32: signal.Signal (trim, DAC -> pA,
    standardise), the network (the reference's taiyaki.layers classes; parameters rounded to
    bf16 first, on both sides, as in make_trained), flipflop_remap.flipflop_remap and
    SignalMapping.from_remapping_path / get_read_dictionary -- the body of
    prepare_mapping_funcs.oneread_remap (:25-116), whose import needs ont_fast5_api (stubbed:
    only its names are imported, the stub is never called).

    Pins: (1) Dacs, Reference and the five scalars of the three reads in
    test/data/mapped_signal_file/mapped_remap_samref.hdf5 -- written by the reference through
    ont_fast5_api + h5py from these fast5 files -- equal what is stored here (that file's
    Ref_to_signal came from another, stride-2, model and is not comparable); (2) the multi-read
    file test/data/multireads holds the same samples; (3) shift / scale of readparams.tsv are
    reproduced bit for bit by med_mad of Signal(read).current (bin/generate_per_read_params.py:
    the UNTRIMMED current)."""
    import types
    import warnings
    for m in ('ont_fast5_api', 'ont_fast5_api.conversion_tools', 'ont_fast5_api.fast5_interface',
              'ont_fast5_api.conversion_tools.conversion_utils'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['ont_fast5_api.conversion_tools.conversion_utils'].get_fast5_file_list = None
    sys.modules['ont_fast5_api.fast5_interface'].get_fast5_file = None
    from taiyaki import flipflop_remap as ref_remap
    from taiyaki import signal as ref_signal
    from taiyaki import signal_mapping as ref_sm
    from taiyaki.maths import med_mad as ref_med_mad
    from taiyaki_b200 import hdf5_min, mapped_signal_files
    from taiyaki_b200.helpers import _legacy_rnn_pickles
    src = open(os.path.join(REF, 'taiyaki/flipflop_remap.py')).read()       # numpy 2 shim, see make_remap
    exec(compile(src.replace('m -= move', 'm -= int(move)'), 'flipflop_remap.py', 'exec'), ref_remap.__dict__)
    data = os.path.join(REF, 'test/data')
    with _legacy_rnn_pickles(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net = torch.load(os.path.join(REF, 'models/mGru_flipflop_remapping_model_r9_DNA.checkpoint'),
                         map_location='cpu', weights_only=False)
    assert type(net).__module__ == 'taiyaki.layers'
    net.eval()
    out = {}
    with torch.no_grad():
        for name, p in net.state_dict().items():
            r = p.detach().to(torch.bfloat16)
            p.copy_(r.float())
            out['param_' + name] = r.view(torch.int16).numpy().view(np.uint16)
    stride = int(net.sublayers[0].stride)
    out['model_cfg'] = np.array([net.sublayers[0].size, stride, net.sublayers[0].winlen])
    params = {}
    for line in open(os.path.join(data, 'readparams.tsv')).read().strip().splitlines()[1:]:
        uuid, t0, t1, shift, scale = line.split('\t')
        params[uuid] = {'trim_start': int(t0), 'trim_end': int(t1), 'shift': float(shift), 'scale': float(scale)}
    refs, name = {}, None
    for line in open(os.path.join(data, 'per_read_references.fasta')):
        if line.startswith('>'):
            name = line[1:].split()[0]
            refs[name] = ''
        elif name is not None:
            refs[name] += line.strip()
    stored = {}
    with mapped_signal_files.HDF5Reader(os.path.join(data, 'mapped_signal_file/mapped_remap_samref.hdf5')) as msr:
        for r in msr.reads():
            stored[r.read_id] = r
    multi = hdf5_min.File(os.path.join(data, 'multireads/FAK40126_2fd3b110ca0d020049836a61f0dfb2b9983808f9_0.fast5'))
    ids = []
    for fn in sorted(os.listdir(os.path.join(data, 'reads'))):
        f = hdf5_min.File(os.path.join(data, 'reads', fn))
        group = sorted(f['Raw/Reads'].keys())[-1]
        attrs = f['Raw/Reads/' + group].attrs
        rid = attrs['read_id'].decode()
        assert fn == rid + '.fast5'
        dacs = f['Raw/Reads/' + group + '/Signal'].read()
        channel = f['UniqueGlobalKey/channel_id'].attrs
        np.testing.assert_array_equal(dacs, multi['read_' + rid + '/Raw/Signal'].read())
        ids.append(rid)
        out[rid + '_dacs'] = dacs
        out[rid + '_read_group'] = np.array(group)
        out[rid + '_channel'] = np.array([channel[k] for k in ('offset', 'range', 'digitisation', 'sampling_rate')])
        out[rid + '_read_attrs'] = np.array([int(attrs[k]) for k in ('start_time', 'duration', 'read_number', 'start_mux')])
        p = params[rid]
        out[rid + '_params'] = np.array([p['trim_start'], p['trim_end'], p['shift'], p['scale']])
        info = {k: float(channel[k]) for k in ('offset', 'range', 'digitisation', 'sampling_rate')}
        whole = ref_signal.Signal(dacs=dacs, channel_info=info, read_id=rid)
        assert tuple(ref_med_mad(whole.current)) == (p['shift'], p['scale']), rid
        if rid not in refs:
            continue
        sig = ref_signal.Signal(dacs=dacs, channel_info=info, read_id=rid, read_params=p)
        x = torch.tensor(sig.standardized_current[:, None, None].astype(np.float32))
        with torch.no_grad():
            trans = net(x).numpy()
        score, path = ref_remap.flipflop_remap(np.squeeze(trans).astype('f8'), refs[rid], alphabet='ACGT',
                                               localpen=0.0)
        int_ref = ref_sm.SignalMapping.get_integer_reference(refs[rid], 'ACGT')
        d = ref_sm.SignalMapping.from_remapping_path(path, int_ref, stride, sig).get_read_dictionary()
        g = stored[rid]
        np.testing.assert_array_equal(d['Dacs'], g.Dacs)
        np.testing.assert_array_equal(d['Reference'], g.Reference)
        assert [d[k] for k in ('shift_frompA', 'scale_frompA', 'range', 'offset', 'digitisation')] == [
            g.shift_frompA, g.scale_frompA, g.range, g.offset, g.digitisation]
        out[rid + '_reference'] = np.array(refs[rid])
        out[rid + '_Ref_to_signal'] = np.asarray(d['Ref_to_signal'], dtype=np.int32)
        out[rid + '_Reference'] = np.asarray(d['Reference'], dtype=np.int16)
        out[rid + '_remap_score'] = np.float64(score)
        out[rid + '_path'] = path.astype(np.int32)
    out['read_ids'] = np.array(ids)
    np.savez_compressed(os.path.join(HERE, 'prepare_remap.npz'), **out)
    print('prepare_remap.npz', os.path.getsize(os.path.join(HERE, 'prepare_remap.npz')), 'bytes;',
          {r[:8]: (len(out[r + '_dacs']), float(out[r + '_remap_score']) if r + '_remap_score' in out else None)
           for r in ids})


def make_model_json():
    """Guppy-format JSON of the reference's shipped checkpoints -> tests/golden/model_json.npz (md5 and
    length of the text only).  Each checkpoint is unpickled into the REFERENCE's taiyaki.layers classes
    (as in make_trained) and dumped the way bin/dump_json.py:24-34 does: `model.json()` plus the file's
    md5sum, `json.dump(..., indent=4, cls=taiyaki.json.JsonEncoder)` for the remapping model, and
    compact (no indent) for the larger mLstm r9.4.1 model."""
    import hashlib
    import json
    import warnings
    from taiyaki.json import JsonEncoder
    from taiyaki_b200.helpers import _legacy_rnn_pickles
    out = {}
    for name, indent in (('mGru_flipflop_remapping_model_r9_DNA', 4), ('mLstm_flipflop_model_r941_DNA', None)):
        fn = os.path.join(REF, 'models', name + '.checkpoint')
        with _legacy_rnn_pickles(), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            net = torch.load(fn, map_location='cpu', weights_only=False)
        assert type(net).__module__ == 'taiyaki.layers'
        json_out = net.json()
        # helpers.file_md5 (taiyaki/helpers.py:302-317; the module needs `imp`, gone in 3.12): md5 of the file
        json_out['md5sum'] = hashlib.md5(open(fn, 'rb').read()).hexdigest()
        text = json.dumps(json_out, indent=indent, cls=JsonEncoder)
        out[name] = np.array([hashlib.md5(text.encode()).hexdigest(), str(len(text)), json_out['md5sum']])
        print(name, out[name])
    np.savez_compressed(os.path.join(HERE, 'model_json.npz'), **out)


MOD_WEIGHT_ALPHABETS = [('ACGTZ', 'ACGTC', ['5mC']), ('ACGTZY', 'ACGTCA', ['5mC', '6mA']),
                        ('ACGTZYX', 'ACGTCAC', ['5mC', '6mA', '5hmC'])]


def mod_weight_reads(nlabel, seed=11, nreads=7):
    """The label sequences both sides of tests/test_host.py::test_mod_prior_weights use."""
    rng = np.random.RandomState(seed)
    return [rng.randint(0, nlabel, size=200 + 13 * i).astype(np.int16) for i in range(nreads)]


def make_mod_weights():
    """AlphabetInfo.compute_log_odds_weights / compute_mod_inv_freq_weights of the reference
    (taiyaki/alphabet.py:35-100) on seeded label sequences, all reads sampled."""
    from taiyaki import alphabet as ref_alphabet

    class Read:
        def __init__(self, ref):
            self.Reference = ref
    out = {}
    for alpha, collapse, names in MOD_WEIGHT_ALPHABETS:
        info = ref_alphabet.AlphabetInfo(alpha, collapse, names)
        refs = mod_weight_reads(len(alpha))
        out[alpha + '_log_odds'] = info.compute_log_odds_weights([Read(r) for r in refs], 100)
        out[alpha + '_inv_freq'] = info.compute_mod_inv_freq_weights([{'Reference': r} for r in refs], 100)
    np.savez_compressed(os.path.join(HERE, 'mod_weights.npz'), **out)
    print('mod_weights.npz:', {k: v.tolist() for k, v in out.items()})


if __name__ == '__main__':
    # `make_golden.py decode` / `make_golden.py basecall` regenerate that file only
    if sys.argv[1:] == ['basecall']:
        make_basecall()
    elif sys.argv[1:] == ['real']:
        make_real()
    elif sys.argv[1:] == ['remap']:
        make_remap()
    elif sys.argv[1:] == ['trained']:
        make_trained()
    elif sys.argv[1:] == ['mod_weights']:
        make_mod_weights()
    elif sys.argv[1:] == ['prepare']:
        make_prepare()
    elif sys.argv[1:] == ['model_json']:
        make_model_json()
    else:
        if sys.argv[1:] != ['decode']:
            main()
        make_decode()
        if sys.argv[1:] != ['decode']:
            make_basecall()
            make_real()
            make_remap()
            make_mod_weights()
            make_prepare()
            make_model_json()
