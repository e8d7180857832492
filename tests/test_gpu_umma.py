"""Known-answer tests of the tcgen05/TMEM engine (csrc/umma.cuh) through the C ABI."""
import pytest
import torch

from cfpnet_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K,shift", [(32, 16, 0), (64, 64, 0), (128, 128, 0), (256, 256, 0), (64, 64, 1),
                                        (128, 64, 37), (32, 32, 138), (16, 16, 0), (48, 32, 3)])
def test_umma_known_answer(N, K, shift):
    g = torch.Generator().manual_seed(N * 1000 + K + shift)
    rows = 128 + shift
    A = torch.randn(rows, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    D = torch.empty(128, N, device="cuda", dtype=torch.float32)
    _lib.call("cfp_selftest_umma", A.data_ptr(), B.data_ptr(), D.data_ptr(), rows, N, K, shift, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = A[shift:shift + 128].double() @ B.double().t()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err          # bf16 products are exact in fp32; only the summation order differs


def test_umma_integer_exact():
    """Small-integer operands: every product and partial sum is exactly representable -> bit-exact."""
    g = torch.Generator().manual_seed(7)
    A = torch.randint(-4, 5, (128, 64), generator=g).to(torch.bfloat16).cuda()
    B = torch.randint(-4, 5, (64, 64), generator=g).to(torch.bfloat16).cuda()
    D = torch.empty(128, 64, device="cuda", dtype=torch.float32)
    _lib.call("cfp_selftest_umma", A.data_ptr(), B.data_ptr(), D.data_ptr(), 128, 64, 64, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(D, A.float() @ B.float().t())
