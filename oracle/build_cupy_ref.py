#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE.  Compiles the reference's GPU path for the partition function
-- the CUDA C sources inside its CuPy RawKernels (taiyaki/cupy_extensions/flipflop.py:10-84
flipflop_fwd, :128-207 flipflop_bwd, :248-295 flipflop_make_trans, :387-466 flipflop_viterbi) --
with nvcc for sm_100a, WHERE THEY LIE: the kernel strings are read out of the reference file at
build time, written with generated launchers into oracle/_ref/ (git-ignored, never committed)
and linked into oracle/_ref/libcupy_ref.so.  CuPy itself is not in this image; the launch
geometry of the generated launchers is the one of the reference's Python wrappers
(flipflop.py:118-124 grid (N) x block (nbase); :241-244 and :330-331 block (2 nbase);
:497-499).  Used by tests/ (parity of csrc/logz.cu against the reference's own kernels on the
same GPU) and by tools/microbench.py (the "reference GPU path" timing); never by the product.

usage: python oracle/build_cupy_ref.py [REFERENCE_ROOT]      (skips quietly when absent)"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')

LAUNCHERS = r'''
// ---- generated launchers (oracle/build_cupy_ref.py); geometry of the reference's wrappers ----
#include <cuda_runtime.h>
extern "C" int ref_cupy_flipflop_fwd(const float *scores, float *fwd, float *fact, long long T, long long N,
                                     long long nbase, void *stream) {
    flipflop_fwd<<<dim3((unsigned)N), dim3((unsigned)nbase), 0, (cudaStream_t)stream>>>(scores, fwd, fact, T, N, nbase);
    return (int)cudaGetLastError();
}
extern "C" int ref_cupy_flipflop_bwd(const float *scores, float *bwd, float *fact, long long T, long long N,
                                     long long nbase, void *stream) {
    flipflop_bwd<<<dim3((unsigned)N), dim3((unsigned)(2 * nbase)), 0, (cudaStream_t)stream>>>(scores, bwd, fact, T, N, nbase);
    return (int)cudaGetLastError();
}
extern "C" int ref_cupy_flipflop_make_trans(const float *scores, const float *fwd, const float *bwd, float *trans,
                                            long long T, long long N, long long nbase, void *stream) {
    flipflop_make_trans<<<dim3((unsigned)N), dim3((unsigned)(2 * nbase)), 0, (cudaStream_t)stream>>>(
        scores, fwd, bwd, trans, T, N, nbase);
    return (int)cudaGetLastError();
}
'''


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    src = os.path.join(ref, 'taiyaki', 'cupy_extensions', 'flipflop.py')
    if not os.path.exists(src):
        print('reference CuPy sources not present; using prebuilt oracle/_ref/libcupy_ref.so if any')
        return 0
    text = open(src).read()
    kernels = re.findall(r"cp\.RawKernel\(r'''(.*?)''',\s*'(\w+)'\)", text, flags=re.S)
    names = [n for _, n in kernels]
    for need in ('flipflop_fwd', 'flipflop_bwd', 'flipflop_make_trans'):
        assert need in names, 'kernel %s not found in %s' % (need, src)
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, 'libcupy_ref.so')
    if os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(src), os.path.getmtime(__file__)):
        return 0
    cu = os.path.join(OUT, 'cupy_kernels.cu')
    with open(cu, 'w') as f:
        f.write('// extracted at build time from %s -- do not commit\n' % src)
        for body, name in kernels:
            if name == 'flipflop_viterbi':
                continue          # needs no launcher here (csrc/viterbi.cu is pinned by golden vectors)
            f.write(body + '\n')
        f.write(LAUNCHERS)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-w', '-Xcompiler', '-fPIC',
                    '-shared', '-o', lib, cu], check=True)
    os.remove(cu)                # the extracted reference text does not stay around, only the binary
    print('built', lib)
    return 0


if __name__ == '__main__':
    sys.exit(main())
