import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pearl_b200 import _lib
lib = _lib.init(0)
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.randn((128, 64), generator=g, device="cuda"); B = torch.randn((64, 64), generator=g, device="cuda")
want = A.double() @ B.double().T; scale = A.abs().double() @ B.abs().double().T
for reps in (1, -1):
    d = torch.zeros((128, 64), device="cuda")
    _lib.check(lib.prl_test_umma_gemm_ts(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(d.data_ptr()), 64, 64, reps, None))
    torch.cuda.synchronize()
    print("reps", reps, "err", ((d.double() - want).abs() / scale).max().item())
