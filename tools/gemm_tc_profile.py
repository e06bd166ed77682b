"""Developer tool: per-chunk SM-clock stamps of k_gemm_tc (CTA 0: loader warp 0 and the issuer warp).
python tools/gemm_tc_profile.py [op M N K engine]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pearl_b200 import _lib  # noqa: E402


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def main():
    op, M, N, K, engine = (int(x) for x in sys.argv[1:6]) if len(sys.argv) >= 6 else (0, 512, 256, 376, 64)
    lib = _lib.init(0)
    dev = "cuda"
    a = torch.randn((M, K if op == 0 else N), device=dev)
    b = torch.randn((N, K), device=dev) if op != 2 else torch.randn((M, K), device=dev)
    bias = torch.randn((N,), device=dev) if op == 0 else None
    c = torch.empty((M, N) if op == 0 else ((M, K) if op == 1 else (N, K)), device=dev)
    ct = torch.empty((N,), device=dev) if op == 2 else None
    st = torch.zeros((33, 8), dtype=torch.int64, device=dev)

    def run():
        _lib.check(lib.prl_test_contraction(op, engine, M, N, K, p(a), p(b), None, 0, p(bias), None, 0, 0, p(c), p(ct), 1, None))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    _lib.check(lib.prl_test_contraction_stamps(p(st)))
    run()
    torch.cuda.synchronize()
    _lib.check(lib.prl_test_contraction_stamps(None))
    s = st.cpu()
    nch = (({0: K, 1: N, 2: M}[op]) + 31) // 32
    t0 = int(s[32, 2])
    print(f"op {op} {M}x{N}x{K} engine {engine}: {nch} chunks; first sync at +{int(s[32, 3]) - t0}")
    print("chunk  iter_start  loads_issued  stage_free  stored  fenced | issuer: operands_ready  mmas_issued")
    for ch in range(min(nch, 32)):
        r = [int(x) - t0 for x in s[ch]]
        print(f"{ch:5d}  {r[0]:10d}  {r[1]:12d}  {r[2]:10d}  {r[4]:6d}  {r[3]:6d} | {r[5]:23d}  {r[6]:11d}")
    print(f"all MMAs done at +{int(s[32, 0]) - t0}, epilogue done at +{int(s[32, 1]) - t0}")


if __name__ == "__main__":
    main()
