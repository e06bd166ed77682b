"""tcgen05 / TMEM building block (pearl_b200/csrc/umma.cuh) against an fp64 product."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(64, 64), (64, 128), (128, 64), (16, 8), (256, 32)])
def test_umma_3xtf32_matches_fp64(n, k):
    from pearl_b200 import _lib
    lib = _lib.init(0)
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + k)
    a = torch.randn((128, k), generator=g, device="cuda")
    b = torch.randn((n, k), generator=g, device="cuda")
    want = (a.double() @ b.double().T)
    scale = (a.abs().double() @ b.abs().double().T)  # sum |a||b|: the natural error scale of a dot product
    for passes, tol in ((3, 4e-6), (1, 2e-3)):
        d = torch.full((128, n), float("nan"), device="cuda")
        _lib.check(lib.prl_test_umma_gemm(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(d.data_ptr()),
                                          n, k, passes, None))
        torch.cuda.synchronize()
        err = ((d.double() - want).abs() / scale).max().item()
        print(f"    N={n} K={k} passes={passes}: max |err| / sum|a||b| = {err:.3e}")
        assert err < tol
        if passes == 1:
            assert err > 1e-5  # really TF32 (a fp32 product would be ~1e-7)
