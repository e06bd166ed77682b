import sys, ctypes as C, torch
sys.path.insert(0, "/root/repo")
from pearl_b200 import _lib
lib = _lib.init(0)
for (n, k) in [(64, 64), (128, 64)]:
    A = torch.randn((128, k), device="cuda"); B = torch.randn((n, k), device="cuda"); d = torch.zeros((128, n), device="cuda")
    _lib.check(lib.prl_test_umma_gemm_ts(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(d.data_ptr()), n, k, 32, None))
    torch.cuda.synchronize()
