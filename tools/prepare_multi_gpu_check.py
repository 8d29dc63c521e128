#!/usr/bin/env python
"""bin/prepare_mapped_reads.py under torchrun on G GPUs, checked against the reference's mappings:
writes the golden fixture reads (tests/golden/prepare_remap.npz) as fast5 files plus parameter
table, references and the shipped remapping model's checkpoint into a scratch directory, launches
`python -m torch.distributed.run --nproc-per-node G bin/prepare_mapped_reads.py ...`, and compares
the joined output with the mappings of the reference's own CPU flow.  Prints one JSON line.

    gpurun --gpus 2 -- 'python tools/prepare_multi_gpu_check.py 2'
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import fast5_fixture  # noqa: E402
from test_fast5 import MAPPED, remapping_model  # noqa: E402
from taiyaki_b200 import helpers, mapped_signal_files  # noqa: E402


def main():
    ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    import torch
    g = fast5_fixture.golden()
    work = tempfile.mkdtemp(prefix='prepare_multi_gpu_')
    reads_dir, tsv, fasta = fast5_fixture.write_inputs(work, g, multi=False)
    ckpt, _ = helpers.save_model(remapping_model(g, torch.device('cpu')), work)
    out = os.path.join(work, 'mapped.hdf5')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(ngpu),
           '--master-addr', '127.0.0.1', '--master-port', '29533',
           os.path.join(ROOT, 'bin', 'prepare_mapped_reads.py'), reads_dir, tsv, out, ckpt, fasta]
    t0 = time.time()
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    seconds = time.time() - t0
    if run.returncode != 0:
        print(run.stdout[-2000:], run.stderr[-4000:], file=sys.stderr)
        raise SystemExit(run.returncode)
    left = sorted(f for f in os.listdir(work) if 'shard' in f)
    agreement = {}
    with mapped_signal_files.MappedSignalReader(out) as msr:
        ids = sorted(msr.get_read_ids())
        check = msr.check()
        for read in msr.reads():
            want = g[read.read_id + '_Ref_to_signal']
            agreement[read.read_id[:8]] = float((np.asarray(read.Ref_to_signal) == want).mean())
            assert np.array_equal(read.Dacs, g[read.read_id + '_dacs'])
    ok = ids == MAPPED and check == 'pass' and not left and min(agreement.values()) >= 0.98
    print(json.dumps({'what': 'prepare_mapped_reads under torchrun', 'n_gpus': ngpu, 'ok': ok, 'reads': len(ids),
                      'check': check, 'shard_files_left': left, 'wall_s': round(seconds, 1),
                      'ref_to_signal_identical_to_reference_flow': agreement,
                      'joined': [line for line in run.stderr.splitlines() if 'joined' in line]}))
    raise SystemExit(0 if ok else 1)


if __name__ == '__main__':
    main()
