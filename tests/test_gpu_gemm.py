"""tcgen05 GEMM (csrc/gemm_tc5.cu) against torch.matmul on the same bf16 operands.
Both accumulate in fp32; differences are summation order only -> 1e-5 relative to
the row's operand magnitude.  Shapes: the contractions of the train step at
config A and their edge cases (K tail, N < tile, M tail, split-K, row map)."""
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


def gemm(A, a_mn, B, b_mn, M, N, K, epi=0, bias=None, scale=1.0, k_splits=1, map_g=0, map_h=0,
         out=None):
    from taiyaki_b200 import _lib
    lib = _lib.lib()
    dev = A.device
    C = torch.full((M, N), float('nan'), device=dev) if out is None else out
    rc = lib.ty_gemm_bf16(_lib.ptr(A), A.stride(0), a_mn, _lib.ptr(B), B.stride(0), b_mn, M, N, K,
                          _lib.ptr(C), C.stride(0), epi, _lib.ptr(bias), scale, k_splits, map_g,
                          map_h, _lib.stream_ptr(dev))
    _lib.check(rc, 'ty_gemm_bf16')
    torch.cuda.synchronize()
    return C


def check(C, ref, K):
    err = (C - ref).abs().max().item()
    mag = ref.abs().max().item()
    assert np.isfinite(err) and err <= 2e-5 * max(mag, 1.0) * max(1.0, (K / 256) ** 0.5), (err, mag)


@pytest.mark.parametrize('M,N,K', [(51200, 1024, 256), (1000, 256, 312), (128, 128, 64),
                                   (200, 40, 256), (333, 200, 1000), (64, 64, 16)])
def test_nt_store(dev, M, N, K):
    torch.manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.randn(M, Kp, device=dev).to(torch.bfloat16)[:, :K]
    B = torch.randn(N, Kp, device=dev).to(torch.bfloat16)[:, :K]
    Np = (N + 3) // 4 * 4
    out = torch.full((M, Np), float('nan'), device=dev)[:, :N]
    C = gemm(A, 0, B, 0, M, N, K, out=out)
    check(C, A.float() @ B.float().t(), K)


def test_bias_tanh_scores(dev):
    """The score projection: [nblk*N, 256] x [40, 256]^T + b -> 5 tanh (layers.py:1411)."""
    torch.manual_seed(1)
    M, N, K = 51200, 40, 256
    A = (0.3 * torch.randn(M, K, device=dev)).to(torch.bfloat16)
    B = (0.3 * torch.randn(N, K, device=dev)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    C = gemm(A, 0, B, 0, M, N, K, epi=1, bias=bias, scale=5.0)
    ref = 5.0 * torch.tanh(A.float() @ B.float().t() + bias)
    assert (C - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize('M,N,K', [(51200, 256, 1024), (300, 72, 200)])
def test_dgrad_b_mn_major(dev, M, N, K):
    """dx = dY W with W stored [K][N] (N contiguous): MN-major B."""
    torch.manual_seed(2)
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = torch.randn(K, N, device=dev).to(torch.bfloat16)
    C = gemm(A, 0, W, 1, M, N, K)
    check(C, A.float() @ W.float(), K)


@pytest.mark.parametrize('M,N,K,splits', [(1024, 256, 51200, 9), (768, 256, 12800, 4),
                                          (40, 256, 5000, 3), (1024, 256, 640, 1)])
def test_wgrad_both_mn_major_split_k(dev, M, N, K, splits):
    """dW = dY^T X: both operands stored [K][.]; split-K accumulation by red.global.add
    into an existing buffer, with and without the unit-major -> gate-major row map."""
    torch.manual_seed(3)
    dY = torch.randn(K, M, device=dev).to(torch.bfloat16)
    X = torch.randn(K, N, device=dev).to(torch.bfloat16)
    ref = dY.float().t() @ X.float()
    base = torch.randn(M, N, device=dev)
    C = gemm(dY, 1, X, 1, M, N, K, epi=2, k_splits=splits, out=base.clone())
    check(C - base, ref, K)
    C = gemm(dY, 1, X, 1, M, N, K, epi=3, k_splits=splits, out=base.clone())
    check(C - base, ref, K)
    if M % 4 == 0:
        G, H = 4, M // 4
        C = gemm(dY, 1, X, 1, M, N, K, epi=2, k_splits=splits, map_g=G, map_h=H,
                 out=torch.zeros(M, N, device=dev))
        # row r = u*G + g of the product lands in row g*H + u
        want = ref.view(H, G, N).transpose(0, 1).reshape(M, N)
        check(C, want, K)


def test_sliced_operands(dev):
    """Views with an offset and a pitch, as the recurrent layers pass them."""
    torch.manual_seed(4)
    T, Nb, H = 50, 8, 256
    d = torch.randn(T, Nb, 4 * H, device=dev).to(torch.bfloat16)
    y = torch.randn(T, Nb, H, device=dev).to(torch.bfloat16)
    a = d[1:].reshape(-1, 4 * H)
    b = y[:-1].reshape(-1, H)
    C = gemm(a, 1, b, 1, 4 * H, H, (T - 1) * Nb, epi=2, k_splits=2, out=torch.zeros(4 * H, H, device=dev))
    check(C, a.float().t() @ b.float(), (T - 1) * Nb)
