"""GPU: the two GEMM cores behind sat_linear (SIMT FFMA fp32/bf16 and tcgen05 bf16) against torch fp32 matmul
on the same (bf16-rounded) operands.  Shapes cover M/N/K tails, ring wrap-around (K > 4 stages * 64) and strided A."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def sat_linear(A, W, bias, c_f32, use_tc, lda=None):
    from sat_b200 import _lib
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=torch.float32 if c_f32 else A.dtype, device=A.device)
    _lib.check(_lib.lib().sat_linear(_lib.ptr(A), lda or A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(bias), _lib.ptr(out), N,
                                     M, N, K, _lib.dtype_code(A.dtype), 1 if c_f32 else 0, 1 if use_tc else 0,
                                     _lib.stream_ptr()), "sat_linear")
    torch.cuda.synchronize()
    return out


SHAPES = [(128, 64, 64), (128, 128, 256), (256, 2688, 512), (200, 136, 72), (8, 6400, 256), (5120, 6400, 256),
          (300, 512, 2048), (1, 8, 8), (129, 72, 1032)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("use_tc", [False, True])
def test_linear_bf16(M, N, K, use_tc):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ W.float().t() + bias
    out = sat_linear(A, W, bias, True, use_tc)
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err          # fp32 accumulation of exact bf16 products
    out16 = sat_linear(A, W, None, False, use_tc)
    ref16 = (A.float() @ W.float().t())
    assert float((out16.float() - ref16).abs().max() / ref16.abs().max()) < 1e-2


@pytest.mark.parametrize("M,N,K", SHAPES[:6])
def test_linear_fp32(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    ref = (A.double() @ W.double().t()).float()
    out = sat_linear(A, W, None, True, False)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-5


def test_linear_tc_strided_a():
    g = torch.Generator(device="cuda").manual_seed(2)
    big = torch.randn(256, 3 * 512, device="cuda", generator=g).to(torch.bfloat16)
    A = big[:, 512:1024]                       # row stride 1536, 16B-aligned offset
    W = (torch.randn(640, 512, device="cuda", generator=g) / 512 ** 0.5).to(torch.bfloat16)
    ref = A.float() @ W.float().t()
    out = sat_linear(A, W, None, True, True, lda=big.stride(0))
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-5
