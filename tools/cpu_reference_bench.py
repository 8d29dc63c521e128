#!/usr/bin/env python
"""Kernel-level CPU reference timings on the host cores of the GPU box (BASELINE.md plan C1-C3):
  C1  the reference's own C (oracle/_ref/libctc_ref.so: crf_flipflop_grad, cat_mod_flipflop_grad,
      OpenMP over chunks) on the same synthetic (scores, sequences) the GPU microbenchmarks use;
  C2  the reference's SPEED_TEST harness (oracle/_ref/crf_speed nbatch nblock ntimes seed);
  C3  the partition function: the C restatement of the oracle (single thread).
This script executes oracle/: it is the CPU-baseline leg, never the product path."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

ncores = os.cpu_count()
os.environ.setdefault('OMP_NUM_THREADS', str(ncores))
os.environ.setdefault('OMP_PROC_BIND', 'true')
out = []


def emit(**kw):
    kw['host_threads'] = int(os.environ['OMP_NUM_THREADS'])
    print(json.dumps(kw))
    out.append(kw)


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


for tag, nblk, S, stride in (('A', 800, 40, 5), ('B', 2000, 45, 2)):
    N = 64
    scores = oracle.synth_scores(nblk, N, S, seed=0)
    seqs, seqlen, raw = oracle.synth_seqs(nblk, N, stride=stride, seed=1)
    alg = 2 * S * 4 * nblk * N
    if not oracle.have_ref():
        emit(what='C1', tag=tag, error='oracle/_ref not built')
        continue
    if S == 40:
        ms = timeit(lambda: oracle.crf_flipflop_loss(scores, seqs, seqlen, 1.0, True, 'ref'))
    else:
        mod_cats = np.concatenate(
            [((r == 1) & (np.random.RandomState(3).uniform(size=len(r)) < 0.5)).astype(np.int64)
             for r in raw])
        ms = timeit(lambda: oracle.cat_mod_flipflop_loss(
            scores, seqs, seqlen, mod_cats, np.array([0, 1, 3, 4, 5]), np.ones(5, np.float32),
            1.0, True, 'ref'))
    emit(what='C1 reference C crf grad (incl. numpy index build)', tag=tag, nblk=nblk, N=N, S=S,
         ms_median=ms, alg_GBps=alg / ms / 1e6)
    w40 = np.ascontiguousarray(scores[:, :, :40])
    ms = timeit(lambda: oracle.c_flipflop_logz(w40, want_grad=True, impl='f32'), reps=3)
    emit(what='C3 logZ + gradient, oracle C restatement (1 thread)', tag=tag, nblk=nblk, N=N,
         ms_median=ms)

speed = os.path.join(ROOT, 'oracle', '_ref', 'crf_speed')
if os.path.exists(speed):
    for nblk in (800, 2000):
        r = subprocess.run([speed, '64', str(nblk), '5', '1'], capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if 'ms' in l or 'Timing' in l or 'took' in l]
        emit(what='C2 SPEED_TEST ./crf_speed 64 %d 5 1' % nblk, stdout=lines[-8:])
else:
    emit(what='C2', error='oracle/_ref/crf_speed not built')
# the same harness shapes on the GPU (scores ~ U(-5, 5), L_b = nblock (1 + (b - N/2) / (5 N)) / 2,
# c_crf_flipflop.c:802-833), when one is present
try:
    import torch
    if torch.cuda.is_available():
        from taiyaki_b200 import ctc
        dev = torch.device('cuda:0')
        for nblk in (800, 2000):
            N = 64
            rng = np.random.RandomState(1)
            sc = torch.tensor((10.0 * rng.uniform(size=(nblk, N, 40)) - 5.0).astype(np.float32),
                              device=dev)
            lens = [int(nblk * (1 + (b - 0.5 * N) / (5.0 * N)) / 2) for b in range(N)]
            seqs, seqlen, _ = oracle.synth_seqs(nblk, N, seed=2, lengths=lens)
            st, sl = torch.tensor(seqs), torch.tensor(seqlen)
            ts = []
            for i in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctc.crf_flipflop_cost_grad(sc, st, sl, 1.0, True)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            emit(what='GPU crf grad on the SPEED_TEST shapes (64 x %d, L ~ nblock / 2)' % nblk,
                 ms_median=float(np.median(ts[3:])))
except Exception as e:       # informational
    emit(what='GPU on SPEED_TEST shapes', error=str(e)[:200])
with open(os.path.join(ROOT, 'gpurun_out', 'cpu_reference_kernels.jsonl'), 'w') as f:
    for o in out:
        f.write(json.dumps(o) + '\n')
