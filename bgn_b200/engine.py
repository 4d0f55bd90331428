"""Thin object wrapper around the C-ABI context (include/bgn_b200.h).

Buffers may be `bytes`/`bytearray`, numpy uint8 arrays (host) or torch uint8
tensors (host, pinned host, or CUDA on the context's device).  Results are
returned as numpy arrays unless an input was a CUDA tensor, in which case the
result is a CUDA tensor too (nothing crosses PCIe in that case).  All compute
happens inside libbgn_b200.so on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _cabi


class BgnError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__("bgn_b200 status %d: %s" % (status, msg))
        self.status = status


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _as_buf(x):
    """-> (address, nbytes, keepalive, is_cuda)"""
    if x is None:
        return None, 0, None, False
    if _is_torch(x):
        assert x.is_contiguous(), "tensor must be contiguous"
        if x.is_cuda:
            # The library works on its own (blocking) stream, which is ordered after the legacy default
            # stream only (include/bgn_b200.h).  A tensor produced on any other torch stream -- a side
            # stream, or per-thread default streams -- must be complete before its pointer is handed over.
            import torch
            cur = torch.cuda.current_stream(x.device)
            if cur.cuda_stream != 0:
                cur.synchronize()
        return x.data_ptr(), x.numel() * x.element_size(), x, x.is_cuda
    if isinstance(x, (bytes, bytearray, memoryview)):
        a = np.frombuffer(x, dtype=np.uint8)
        return a.ctypes.data, a.nbytes, (a, x), False
    a = np.ascontiguousarray(x)
    return a.ctypes.data, a.nbytes, a, False


class DeviceBatch:
    """A batch of group elements resident on the engine's GPU in the kernels' own form (bgn_buf,
    include/bgn_b200.h): what the `*_h` entry points consume and produce.  `kind`: 1 = G1 (level 1),
    2 = GT (level 2)."""

    def __init__(self, engine: "Engine", handle: C.c_void_p):
        self.engine, self._h = engine, handle

    @property
    def kind(self) -> int:
        k, n = C.c_int(), C.c_size_t()
        self.engine._lib.bgn_buf_info(self._h, C.byref(k), C.byref(n))
        return k.value

    def __len__(self) -> int:
        k, n = C.c_int(), C.c_size_t()
        self.engine._lib.bgn_buf_info(self._h, C.byref(k), C.byref(n))
        return n.value

    def to_bytes(self, out=None):
        """Element.Bytes() of every element (PBC format); `out` may be a CUDA tensor"""
        o = out if out is not None else np.empty(len(self) * self.engine.elem_bytes, dtype=np.uint8)
        self.engine._check(self.engine._lib.bgn_buf_export(self.engine._ctx, self._h, _as_buf(o)[0]))
        return o

    def free(self):
        if self._h is not None and self._h.value:
            self.engine._lib.bgn_buf_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.engine._ctx.value:
                self.free()
        except Exception:
            pass


class Engine:
    """One BGN key resident on one GPU (bgn_ctx)."""

    def __init__(self, p: int, n: int, l: int, P_bytes: bytes, Q_bytes: bytes, device: int = 0):
        self._lib = _cabi.load()
        self._ctx = C.c_void_p()
        pb = p.to_bytes((p.bit_length() + 7) // 8, "big")
        nb = n.to_bytes((n.bit_length() + 7) // 8, "big")
        prm = _cabi.bgn_params(pb, len(pb), nb, len(nb), l, bytes(P_bytes), bytes(Q_bytes))
        st = self._lib.bgn_ctx_create(C.byref(prm), device, C.byref(self._ctx))
        if st != 0:
            self._ctx = C.c_void_p()
            raise BgnError(st, self._lib.bgn_global_last_error().decode() or "bgn_ctx_create failed")
        L, B, nbytes = C.c_int(), C.c_int(), C.c_int()
        self._lib.bgn_ctx_info(self._ctx, C.byref(L), C.byref(B), C.byref(nbytes))
        self.limbs, self.coord_bytes, self.scalar_bytes = L.value, B.value, nbytes.value
        self.elem_bytes = 2 * B.value
        self.device = device
        self.p, self.n, self.l = p, n, l

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.bgn_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, st: int):
        if st != 0:
            raise BgnError(st, self._lib.bgn_last_error(self._ctx).decode())

    def _out(self, nbytes: int, cuda: bool, out=None):
        if out is not None:
            return out
        if cuda:
            import torch
            return torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % self.device)
        return np.empty(nbytes, dtype=np.uint8)

    def _count(self, nbytes: int) -> int:
        assert nbytes % self.elem_bytes == 0, "buffer is not a whole number of elements"
        return nbytes // self.elem_bytes

    def _binop(self, fn, a, b, out=None):
        pa, na, ka, ca = _as_buf(a)
        pb, nb_, kb, cb = _as_buf(b)
        assert na == nb_
        count = self._count(na)
        o = self._out(na, ca or cb, out)
        po = _as_buf(o)[0]
        self._check(fn(self._ctx, pa, pb, count, po))
        return o

    def _unop(self, fn, a, out=None):
        pa, na, ka, ca = _as_buf(a)
        count = self._count(na)
        o = self._out(na, ca, out)
        self._check(fn(self._ctx, pa, count, _as_buf(o)[0]))
        return o

    def scalars_be(self, ks: Sequence[int], width: Optional[int] = None) -> np.ndarray:
        width = width or self.scalar_bytes
        return np.frombuffer(b"".join(int(k).to_bytes(width, "big") for k in ks), dtype=np.uint8).copy()

    # ------------------------------------------------------------------ C-ABI calls
    def set_option(self, name: str, value: int):
        """tuning knobs (include/bgn_b200.h: bgn_ctx_set_option); none changes any result"""
        self._check(self._lib.bgn_ctx_set_option(self._ctx, name.encode(), int(value)))

    def set_secret(self, q1: int, msg_space: int, baby_steps: int = 0):
        qb = q1.to_bytes((q1.bit_length() + 7) // 8, "big")
        self._check(self._lib.bgn_ctx_set_secret(self._ctx, qb, len(qb), msg_space, baby_steps))

    def encrypt_batch(self, x, r_be=None, out=None):
        """x: int64 array/tensor; r_be: count*scalar_bytes big-endian randomness or None."""
        if not _is_torch(x):
            x = np.ascontiguousarray(x, dtype=np.int64)
        px, nx, kx, cx = _as_buf(x)
        count = nx // 8
        pr, nr, kr, cr = _as_buf(r_be)
        if r_be is not None:
            assert nr == count * self.scalar_bytes
        o = self._out(count * self.elem_bytes, cx or cr, out)
        self._check(self._lib.bgn_encrypt_batch(self._ctx, px, pr, count, _as_buf(o)[0]))
        return o

    def g1_add_batch(self, a, b, out=None):
        return self._binop(self._lib.bgn_g1_add_batch, a, b, out)

    def g1_sub_batch(self, a, b, out=None):
        return self._binop(self._lib.bgn_g1_sub_batch, a, b, out)

    def g1_neg_batch(self, a, out=None):
        return self._unop(self._lib.bgn_g1_neg_batch, a, out)

    def g1_mulconst_batch(self, a, k_be, kbytes: int, out=None):
        pa, na, ka, ca = _as_buf(a)
        pk_, nk, kk, ck = _as_buf(k_be)
        count = self._count(na)
        assert nk == count * kbytes
        o = self._out(na, ca, out)
        self._check(self._lib.bgn_g1_mulconst_batch(self._ctx, pa, pk_, kbytes, count, _as_buf(o)[0]))
        return o

    def gt_mul_batch(self, a, b, out=None):
        return self._binop(self._lib.bgn_gt_mul_batch, a, b, out)

    def gt_div_batch(self, a, b, out=None):
        return self._binop(self._lib.bgn_gt_div_batch, a, b, out)

    def gt_inv_batch(self, a, out=None):
        return self._unop(self._lib.bgn_gt_inv_batch, a, out)

    def gt_pow_batch(self, a, k_be, kbytes: int, out=None):
        pa, na, ka, ca = _as_buf(a)
        pk_, nk, kk, ck = _as_buf(k_be)
        count = self._count(na)
        assert nk == count * kbytes
        o = self._out(na, ca, out)
        self._check(self._lib.bgn_gt_pow_batch(self._ctx, pa, pk_, kbytes, count, _as_buf(o)[0]))
        return o

    def pair_batch(self, a, b, out=None):
        return self._binop(self._lib.bgn_pair_batch, a, b, out)

    def make_l2_batch(self, a, out=None):
        return self._unop(self._lib.bgn_make_l2_batch, a, out)

    def multpoly_batch(self, c1, d1: int, c2, d2: int, count: int, out=None):
        p1, n1, k1, cu1 = _as_buf(c1)
        p2, n2, k2, cu2 = _as_buf(c2)
        assert n1 == count * d1 * self.elem_bytes and n2 == count * d2 * self.elem_bytes
        o = self._out(count * (d1 + d2) * self.elem_bytes, cu1 or cu2, out)
        self._check(self._lib.bgn_multpoly_batch(self._ctx, p1, d1, p2, d2, count, _as_buf(o)[0]))
        return o

    def l2_sum_reduce(self, terms, nterms: int, ncoeff: int, out=None):
        pt, nt, kt, ct = _as_buf(terms)
        assert nt == nterms * ncoeff * self.elem_bytes
        o = self._out(ncoeff * self.elem_bytes, ct, out)
        self._check(self._lib.bgn_l2_sum_reduce(self._ctx, pt, nterms, ncoeff, _as_buf(o)[0]))
        return o

    def _blind(self, fn, a, r_be, out=None):
        pa, na, ka, ca = _as_buf(a)
        pr, nr, kr, cr = _as_buf(r_be)
        count = self._count(na)
        assert nr == count * self.scalar_bytes
        o = self._out(na, ca or cr, out)
        self._check(fn(self._ctx, pa, pr, count, _as_buf(o)[0]))
        return o

    def g1_blind_batch(self, a, r_be, out=None):
        """a[i] + r[i]*Q (level-1 re-randomisation of the non-deterministic mode)."""
        return self._blind(self._lib.bgn_g1_blind_batch, a, r_be, out)

    def gt_blind_batch(self, a, r_be, out=None):
        """a[i] * e(Q,Q)^r[i] (level-2 re-randomisation)."""
        return self._blind(self._lib.bgn_gt_blind_batch, a, r_be, out)

    def multconstpoly_batch(self, cts, d: int, is_l2: bool, digits: Sequence[int], negate: bool, count: int, out=None):
        """MultConstPoly over `count` polynomials of d slots -> count*(d+len(digits)) elements."""
        pc, nc, kc, cc = _as_buf(cts)
        assert nc == count * d * self.elem_bytes
        dg = bytes(bytearray(int(x) for x in digits))
        o = self._out(count * (d + len(dg)) * self.elem_bytes, cc, out)
        self._check(self._lib.bgn_multconstpoly_batch(self._ctx, pc, d, 1 if is_l2 else 0, dg, len(dg),
                                                      1 if negate else 0, count, _as_buf(o)[0]))
        return o

    def evalpoly_batch(self, cts, d: int, is_l2: bool, base: int, count: int, out=None):
        """EvalPoly: one element per polynomial, sum_i base^i * c_i."""
        pc, nc, kc, cc = _as_buf(cts)
        assert nc == count * d * self.elem_bytes
        o = self._out(count * self.elem_bytes, cc, out)
        self._check(self._lib.bgn_evalpoly_batch(self._ctx, pc, d, 1 if is_l2 else 0, base, count, _as_buf(o)[0]))
        return o

    def make_poly_l2_batch(self, cts, d: int, count: int, out=None):
        """MakePolyL2 (deterministic): count*(d+1) GT elements."""
        pc, nc, kc, cc = _as_buf(cts)
        assert nc == count * d * self.elem_bytes
        o = self._out(count * (d + 1) * self.elem_bytes, cc, out)
        self._check(self._lib.bgn_make_poly_l2_batch(self._ctx, pc, d, count, _as_buf(o)[0]))
        return o

    def gt_pow_secret_batch(self, a, out=None):
        return self._unop(self._lib.bgn_gt_pow_secret_batch, a, out)

    def decrypt_batch(self, cts, is_l2: bool):
        """-> (int64 values, uint8 status); status 1 = 'cannot find discrete log; out of bounds'."""
        pc, nc, kc, cc = _as_buf(cts)
        count = self._count(nc)
        if cc:
            import torch
            dev = "cuda:%d" % self.device
            vals = torch.empty(count, dtype=torch.int64, device=dev)
            status = torch.empty(count, dtype=torch.uint8, device=dev)
        else:
            vals = np.empty(count, dtype=np.int64)
            status = np.empty(count, dtype=np.uint8)
        self._check(self._lib.bgn_decrypt_batch(self._ctx, pc, 1 if is_l2 else 0, count, _as_buf(vals)[0],
                                                _as_buf(status)[0]))
        return vals, status

    # ------------------------------------------------------------------ device-resident batches
    def _new_out(self, out: Optional[DeviceBatch]) -> C.c_void_p:
        return out._h if out is not None else C.c_void_p()

    def _wrap(self, h: C.c_void_p, out: Optional[DeviceBatch]) -> DeviceBatch:
        if out is not None:
            out._h = h
            return out
        return DeviceBatch(self, h)

    def import_batch(self, kind: int, data, out: Optional[DeviceBatch] = None) -> DeviceBatch:
        """bytes (PBC format, host or CUDA) -> DeviceBatch; kind 1 = G1 (curve check applies), 2 = GT"""
        pd, nd, kd, cd = _as_buf(data)
        h = self._new_out(out)
        self._check(self._lib.bgn_buf_import(self._ctx, kind, pd, self._count(nd), C.byref(h)))
        return self._wrap(h, out)

    def encrypt_h(self, x, r_be=None, out: Optional[DeviceBatch] = None) -> DeviceBatch:
        if not _is_torch(x):
            x = np.ascontiguousarray(x, dtype=np.int64)
        px, nx, kx, cx = _as_buf(x)
        pr, nr, kr, cr = _as_buf(r_be)
        h = self._new_out(out)
        self._check(self._lib.bgn_encrypt_h(self._ctx, px, pr, nx // 8, C.byref(h)))
        return self._wrap(h, out)

    def g1_add_h(self, a: DeviceBatch, b: DeviceBatch, subtract: bool = False, out: Optional[DeviceBatch] = None):
        h = self._new_out(out)
        self._check(self._lib.bgn_g1_add_h(self._ctx, a._h, b._h, 1 if subtract else 0, C.byref(h)))
        return self._wrap(h, out)

    def gt_mul_h(self, a: DeviceBatch, b: DeviceBatch, divide: bool = False, out: Optional[DeviceBatch] = None):
        h = self._new_out(out)
        self._check(self._lib.bgn_gt_mul_h(self._ctx, a._h, b._h, 1 if divide else 0, C.byref(h)))
        return self._wrap(h, out)

    def pair_h(self, a: DeviceBatch, b: Optional[DeviceBatch] = None, out: Optional[DeviceBatch] = None):
        """e(a[i], b[i]); b = None: makeL2, e(a[i], P)"""
        h = self._new_out(out)
        self._check(self._lib.bgn_pair_h(self._ctx, a._h, b._h if b is not None else None, C.byref(h)))
        return self._wrap(h, out)

    def multpoly_h(self, c1: DeviceBatch, d1: int, c2: DeviceBatch, d2: int, count: int, out: Optional[DeviceBatch] = None):
        h = self._new_out(out)
        self._check(self._lib.bgn_multpoly_h(self._ctx, c1._h, d1, c2._h, d2, count, C.byref(h)))
        return self._wrap(h, out)

    def l2_sum_reduce_h(self, terms: DeviceBatch, nterms: int, ncoeff: int, out: Optional[DeviceBatch] = None):
        h = self._new_out(out)
        self._check(self._lib.bgn_l2_sum_reduce_h(self._ctx, terms._h, nterms, ncoeff, C.byref(h)))
        return self._wrap(h, out)

    def decrypt_h(self, cts: DeviceBatch):
        """-> (int64 values, uint8 status) as numpy arrays; the level is the batch's kind"""
        count = len(cts)
        vals = np.empty(count, dtype=np.int64)
        status = np.empty(count, dtype=np.uint8)
        self._check(self._lib.bgn_decrypt_h(self._ctx, cts._h, _as_buf(vals)[0], _as_buf(status)[0]))
        return vals, status

    # ------------------------------------------------------------------ instrumentation
    def timing_enable(self, on: bool = True):
        self._check(self._lib.bgn_timing_enable(self._ctx, 1 if on else 0))

    def timing_reset(self):
        self._check(self._lib.bgn_timing_reset(self._ctx))

    def timing_get(self, prefix: str = ""):
        ms, n = C.c_double(), C.c_uint64()
        self._check(self._lib.bgn_timing_get(self._ctx, prefix.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timing_last_call(self) -> float:
        """device ms of the last call (copy-in .. copy-out) on the context's stream"""
        ms = C.c_double()
        self._check(self._lib.bgn_timing_last_call(self._ctx, C.byref(ms)))
        return ms.value

    def bench_mulmod(self, ilp: int, iters: int, blocks: int, threads: int) -> float:
        ms = C.c_float()
        self._check(self._lib.bgn_bench_mulmod(self._ctx, ilp, iters, blocks, threads, C.byref(ms)))
        return ms.value


class EngineGroup:
    """One key on SEVERAL GPUs behind one handle (bgn_group, include/bgn_b200.h): every batch call is cut
    into contiguous shards, one per listed device, each run from its own host thread inside the library.
    Buffers are host arrays.  `devices` may repeat an ordinal (its contexts take turns)."""

    def __init__(self, p: int, n: int, l: int, P_bytes: bytes, Q_bytes: bytes, devices: Sequence[int]):
        self._lib = _cabi.load()
        self._grp = C.c_void_p()
        pb = p.to_bytes((p.bit_length() + 7) // 8, "big")
        nb = n.to_bytes((n.bit_length() + 7) // 8, "big")
        prm = _cabi.bgn_params(pb, len(pb), nb, len(nb), l, bytes(P_bytes), bytes(Q_bytes))
        devs = (C.c_int * len(devices))(*devices)
        st = self._lib.bgn_group_create(C.byref(prm), len(devices), devs, C.byref(self._grp))
        if st != 0:
            self._grp = C.c_void_p()
            raise BgnError(st, self._lib.bgn_global_last_error().decode() or "bgn_group_create failed")
        L, B, nbytes = C.c_int(), C.c_int(), C.c_int()
        self._lib.bgn_ctx_info(self._lib.bgn_group_ctx(self._grp, 0), C.byref(L), C.byref(B), C.byref(nbytes))
        self.limbs, self.coord_bytes, self.scalar_bytes = L.value, B.value, nbytes.value
        self.elem_bytes = 2 * B.value
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_grp", None) is not None and self._grp.value:
            self._lib.bgn_group_destroy(self._grp)
            self._grp = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != 0:
            raise BgnError(st, self._lib.bgn_group_last_error(self._grp).decode())

    @staticmethod
    def _host(x, dtype=np.uint8):
        if _is_torch(x):
            assert not x.is_cuda, "group calls take host buffers (a device pointer belongs to one GPU)"
            return x
        return np.ascontiguousarray(x, dtype=dtype)

    def scalars_be(self, ks: Sequence[int], width: Optional[int] = None) -> np.ndarray:
        width = width or self.scalar_bytes
        return np.frombuffer(b"".join(int(k).to_bytes(width, "big") for k in ks), dtype=np.uint8).copy()

    def set_secret(self, q1: int, msg_space: int, baby_steps: int = 0):
        qb = q1.to_bytes((q1.bit_length() + 7) // 8, "big")
        self._check(self._lib.bgn_group_set_secret(self._grp, qb, len(qb), msg_space, baby_steps))

    def set_option(self, name: str, value: int):
        self._check(self._lib.bgn_group_set_option(self._grp, name.encode(), int(value)))

    def encrypt_batch(self, x, r_be=None, out=None):
        x = self._host(x, np.int64)
        count = _as_buf(x)[1] // 8
        r = None if r_be is None else self._host(r_be)
        o = out if out is not None else np.empty(count * self.elem_bytes, dtype=np.uint8)
        self._check(self._lib.bgn_group_encrypt_batch(self._grp, _as_buf(x)[0], _as_buf(r)[0], count, _as_buf(o)[0]))
        return o

    def g1_add_batch(self, a, b, out=None):
        a, b = self._host(a), self._host(b)
        na = _as_buf(a)[1]
        o = out if out is not None else np.empty(na, dtype=np.uint8)
        self._check(self._lib.bgn_group_g1_add_batch(self._grp, _as_buf(a)[0], _as_buf(b)[0], na // self.elem_bytes, _as_buf(o)[0]))
        return o

    def multpoly_batch(self, c1, d1: int, c2, d2: int, count: int, out=None):
        c1, c2 = self._host(c1), self._host(c2)
        o = out if out is not None else np.empty(count * (d1 + d2) * self.elem_bytes, dtype=np.uint8)
        self._check(self._lib.bgn_group_multpoly_batch(self._grp, _as_buf(c1)[0], d1, _as_buf(c2)[0], d2, count, _as_buf(o)[0]))
        return o

    def inner_product(self, c1, d1: int, c2, d2: int, count: int):
        c1, c2 = self._host(c1), self._host(c2)
        o = np.empty((d1 + d2) * self.elem_bytes, dtype=np.uint8)
        self._check(self._lib.bgn_group_inner_product(self._grp, _as_buf(c1)[0], d1, _as_buf(c2)[0], d2, count, _as_buf(o)[0]))
        return o

    def decrypt_batch(self, cts, is_l2: bool):
        cts = self._host(cts)
        count = _as_buf(cts)[1] // self.elem_bytes
        vals = np.empty(count, dtype=np.int64)
        status = np.empty(count, dtype=np.uint8)
        self._check(self._lib.bgn_group_decrypt_batch(self._grp, _as_buf(cts)[0], 1 if is_l2 else 0, count,
                                                      _as_buf(vals)[0], _as_buf(status)[0]))
        return vals, status


def bench_imad_peak(device: int, iters: int, blocks: int, threads: int):
    """-> (ms, IMAD.WIDE instructions per thread)"""
    lib = _cabi.load()
    ms, ipt = C.c_float(), C.c_double()
    st = lib.bgn_bench_imad_peak(device, iters, blocks, threads, C.byref(ms), C.byref(ipt))
    if st != 0:
        raise BgnError(st, "bgn_bench_imad_peak failed")
    return ms.value, ipt.value


def bench_issue_mix(device: int, mix: int, iters: int, blocks: int, threads: int):
    """-> (ms, [IMAD.WIDE, IMAD.LO, IMAD.HI, FFMA, DFMA] instructions per thread)"""
    lib = _cabi.load()
    ms, per = C.c_float(), (C.c_double * 5)()
    st = lib.bgn_bench_issue_mix(device, mix, iters, blocks, threads, C.byref(ms), per)
    if st != 0:
        raise BgnError(st, "bgn_bench_issue_mix failed")
    return ms.value, list(per)
