"""TEST INFRASTRUCTURE. ctypes binding of oracle/libmmoracle.so (the CPU restatement)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmmoracle.so")


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, nthreads=None):
        if not os.path.exists(_LIB):
            build()
        self.L = ctypes.CDLL(_LIB)
        self.nthreads = nthreads or os.cpu_count() or 1
        for f in ("mmo_sinf", "mmo_cosf", "mmo_simplex2", "mmo_simplex3", "mmo_powf", "mmo_rng3_u01"):
            getattr(self.L, f).restype = ctypes.c_float
        self.L.mmo_sinf.argtypes = [ctypes.c_float]
        self.L.mmo_cosf.argtypes = [ctypes.c_float]
        self.L.mmo_powf.argtypes = [ctypes.c_float, ctypes.c_float]
        self.L.mmo_simplex2.argtypes = [ctypes.c_float, ctypes.c_float]
        self.L.mmo_simplex3.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float]
        self.L.mmo_hash.restype = ctypes.c_uint32
        self.L.mmo_hash.argtypes = [ctypes.c_uint32]

    def heightfields(self, origins):
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h = np.empty((n, 256), np.float32)
        w = np.empty((n, 24, 256), np.float32)
        self.L.mmo_heightfields(n, _ptr(origins), _ptr(h), _ptr(w), self.nthreads)
        return h, w
