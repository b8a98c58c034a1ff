"""The tcgen05 3xTF32 blend contraction against fp64 and against the FP32-pipe reference kernel."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(fn, *args):
    from ihmr_b200 import _lib
    _lib.check(fn(*args, C.c_void_p(torch.cuda.current_stream().cuda_stream)), fn.__name__)


@pytest.mark.parametrize("M,Nc,K", [(128, 256, 32), (1, 16, 32), (200, 160, 2336), (333, 2336, 160), (1000, 2336, 160)])
def test_tf32x3_matches_fp64(M, Nc, K):
    from ihmr_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + Nc + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(Nc, K, device="cuda", generator=g) * 0.01        # posedirs-like magnitudes
    Cc = torch.full((M, Nc), float("nan"), device="cuda")
    _run(lib.ihmr_gemm_tf32x3, M, Nc, K, C.c_void_p(A.data_ptr()), K, C.c_void_p(B.data_ptr()), K, C.c_void_p(Cc.data_ptr()), Nc)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    scale = float(ref.abs().max())
    err = float((Cc.double() - ref).abs().max())
    assert np.isfinite(err) and err <= (2e-6 + 3e-8 * (K // 8) * 3) * scale + 1e-9, (err, scale)   # TMEM accumulation rounds toward zero: ~2^-25 per MMA
    # the FP32-pipe kernel (the checker) must agree as well
    Cs = torch.empty(M, Nc, device="cuda")
    Bt = B.T.contiguous()
    _run(lib.ihmr_gemm_reference_fp32, M, Nc, K, C.c_void_p(A.data_ptr()), K, C.c_void_p(Bt.data_ptr()), Nc, C.c_void_p(Cs.data_ptr()), Nc)
    torch.cuda.synchronize()
    assert float((Cs.double() - ref).abs().max()) <= 2e-6 * scale + 1e-9


def test_tf32x3_is_deterministic():
    from ihmr_b200 import _lib
    lib = _lib.load()
    A = torch.randn(500, 160, device="cuda")
    B = torch.randn(2336, 160, device="cuda")
    outs = []
    for _ in range(2):
        Cc = torch.empty(500, 2336, device="cuda")
        _run(lib.ihmr_gemm_tf32x3, 500, 2336, 160, C.c_void_p(A.data_ptr()), 160, C.c_void_p(B.data_ptr()), 160, C.c_void_p(Cc.data_ptr()), 2336)
        outs.append(Cc)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
