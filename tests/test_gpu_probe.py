"""tcgen05 bring-up probe: one 128xNx64 UMMA through the production swizzle/descriptor helpers."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [256, 128, 64])
def test_umma_probe_matches_integer_gemm(N):
    from benerf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(N)
    A = torch.randint(-4, 5, (128, 64), generator=g).half().cuda()
    B = torch.randint(-4, 5, (N, 64), generator=g).half().cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.bnrf_debug_umma_probe(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), N, 0, C.c_void_p(D.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    want = A.float() @ B.float().t()          # small integers: exact in fp16 inputs / fp32 accumulate
    assert torch.equal(D, want), f"max abs diff {(D - want).abs().max().item()}"
