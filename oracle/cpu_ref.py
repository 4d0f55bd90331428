"""ctypes driver for oracle/libcpu_ref.so (the C port of the oracle).  TEST / BASELINE
INFRASTRUCTURE ONLY -- see the header of cpu_ref.c.  Byte-level API mirroring the C-ABI of
the product so tests can compare buffers directly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libcpu_ref.so")
        src = os.path.join(HERE, "cpu_ref.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["make", "-C", HERE])
        # a -march=native build for THIS host's CPU (mulx/adx help the 64-bit-limb Montgomery product by
        # ~30 %), keyed by the CPU flags so a copy built elsewhere is never loaded on a different CPU
        try:
            import hashlib
            with open("/proc/cpuinfo") as f:
                flags = next((ln for ln in f if ln.startswith("flags")), "")
            tag = hashlib.sha1(flags.encode()).hexdigest()[:10]
            nat = os.path.join(HERE, "libcpu_ref_native_%s.so" % tag)
            if not os.path.exists(nat) or os.path.getmtime(src) > os.path.getmtime(nat):
                subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-pthread", "-o", nat, src],
                                      stderr=subprocess.DEVNULL)
            so = nat
        except Exception:
            pass
        L = C.CDLL(so)
        L.cpuref_create.restype = C.c_void_p
        L.cpuref_create.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint64]
        L.cpuref_destroy.argtypes = [C.c_void_p]
        L.cpuref_coord_bytes.argtypes = [C.c_void_p]
        vp = C.c_void_p
        L.cpuref_pair_batch.argtypes = [vp, vp, vp, C.c_size_t, vp, C.c_int]
        L.cpuref_multpoly_batch.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_size_t, vp, C.c_int]
        L.cpuref_encrypt_batch.argtypes = [vp, C.c_char_p, C.c_char_p, vp, vp, C.c_int, C.c_size_t, vp, C.c_int]
        L.cpuref_gt_pow_batch.argtypes = [vp, vp, C.c_char_p, C.c_int, C.c_size_t, vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


class CpuRef:
    def __init__(self, p: int, n: int, l: int, threads: int = 1):
        pb = p.to_bytes((p.bit_length() + 7) // 8, "big")
        nb = n.to_bytes((n.bit_length() + 7) // 8, "big")
        self.h = lib().cpuref_create(pb, len(pb), nb, len(nb), l)
        assert self.h, "unsupported parameters"
        self.B = lib().cpuref_coord_bytes(self.h)
        self.E = 2 * self.B
        self.nbytes = len(nb)
        self.threads = threads

    def __del__(self):
        if getattr(self, "h", None):
            lib().cpuref_destroy(self.h)
            self.h = None

    def pair_batch(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a, b = np.ascontiguousarray(a, dtype=np.uint8), np.ascontiguousarray(b, dtype=np.uint8)
        count = a.size // self.E
        out = np.empty(count * self.E, dtype=np.uint8)
        lib().cpuref_pair_batch(self.h, _p(a), _p(b), count, _p(out), self.threads)
        return out

    def multpoly_batch(self, c1, d1, c2, d2, count) -> np.ndarray:
        c1, c2 = np.ascontiguousarray(c1, dtype=np.uint8), np.ascontiguousarray(c2, dtype=np.uint8)
        out = np.empty(count * (d1 + d2) * self.E, dtype=np.uint8)
        lib().cpuref_multpoly_batch(self.h, _p(c1), d1, _p(c2), d2, count, _p(out), self.threads)
        return out

    def encrypt_batch(self, P: bytes, Q: bytes, x, r_be) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int64)
        r = None if r_be is None else np.ascontiguousarray(r_be, dtype=np.uint8)
        out = np.empty(x.size * self.E, dtype=np.uint8)
        lib().cpuref_encrypt_batch(self.h, P, Q, _p(x), _p(r), self.nbytes, x.size, _p(out), self.threads)
        return out

    def gt_pow_batch(self, a, e: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        count = a.size // self.E
        eb = e.to_bytes(max(1, (e.bit_length() + 7) // 8), "big")
        out = np.empty(a.size, dtype=np.uint8)
        lib().cpuref_gt_pow_batch(self.h, _p(a), eb, len(eb), count, _p(out), self.threads)
        return out
