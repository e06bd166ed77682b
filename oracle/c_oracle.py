"""ctypes loader for oracle/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "mt_sample_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.orc_mt_init_by_array.argtypes = [u32p, u32p, ctypes.c_int]
        L.orc_mt_seed_u64.argtypes = [u32p, ctypes.c_uint64]
        L.orc_mt_genrand_uint32.argtypes = [u32p]
        L.orc_mt_genrand_uint32.restype = ctypes.c_uint32
        L.orc_randbelow.argtypes = [u32p, ctypes.c_uint32]
        L.orc_randbelow.restype = ctypes.c_uint32
        L.orc_sample_setsize.argtypes = [ctypes.c_int64]
        L.orc_sample_setsize.restype = ctypes.c_int64
        L.orc_sample.argtypes = [u32p, ctypes.c_int64, ctypes.c_int64,
                                 ctypes.POINTER(ctypes.c_int64)]
        L.orc_sample.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


class MT:
    """MT19937 state in `random.getstate()[1]` layout (uint32[625])."""

    def __init__(self, state=None, seed=None):
        self.st = np.zeros(625, dtype=np.uint32)
        if state is not None:
            self.st[:] = np.asarray(state, dtype=np.uint32)
        elif seed is not None:
            self.seed(seed)

    def seed(self, seed: int) -> None:
        seed = abs(int(seed))
        key = []
        while True:
            key.append(seed & 0xFFFFFFFF)
            seed >>= 32
            if seed == 0:
                break
        k = np.asarray(key, dtype=np.uint32)
        lib().orc_mt_init_by_array(_p(self.st), _p(k), len(key))

    def getrandbits32(self) -> int:
        return int(lib().orc_mt_genrand_uint32(_p(self.st)))

    def sample(self, n: int, k: int) -> np.ndarray:
        out = np.zeros(max(k, 1), dtype=np.int64)
        rc = lib().orc_sample(_p(self.st), n, k, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        if rc != 0:
            raise ValueError("Sample larger than population or is negative")
        return out[:k]
