"""ctypes driver for tests/hostsim/libhostsim.so: runs the device programs of
bgn_b200/csrc on the CPU (carry flag / grid emulated) so their logic can be
checked against the oracle without a GPU.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
MAXL, MAX_NAF, MAX_EXPW, NSLOT = 34, 1100, 34, 12
SUPPORTED_L = (3, 5, 17)  # limb counts instantiated in hostsim.cpp
PRODUCT_L = (3, 5, 9, 17, 33)  # limb counts the CUDA library instantiates

u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class FieldConsts(C.Structure):
    _fields_ = [("p", C.c_uint32 * MAXL), ("p2", C.c_uint32 * MAXL), ("one", C.c_uint32 * MAXL),
                ("r2", C.c_uint32 * MAXL), ("p4", C.c_uint32 * MAXL), ("p8", C.c_uint32 * MAXL),
                ("p16", C.c_uint32 * MAXL), ("np0", C.c_uint32), ("pad", C.c_uint32 * 3)]


class PairConsts(C.Structure):
    _fields_ = [("l", C.c_uint64), ("naf_len", C.c_int32), ("exp_bits", C.c_int32),
                ("exp", C.c_uint32 * MAX_EXPW), ("naf", C.c_int8 * MAX_NAF),
                ("exp_naf_len", C.c_int32), ("exp_naf", C.c_int8 * MAX_NAF)]


class MillerArgs(C.Structure):
    _fields_ = [("Mx", u32p), ("My", u32p), ("Minf", u8p), ("Ex", u32p), ("Ey", u32p), ("Einf", u8p),
                ("priv", u32p), ("out_re", u32p), ("out_im", u32p), ("NM", C.c_int), ("NE", C.c_int), ("NOUT", C.c_int),
                ("e_bcast", C.c_int), ("dM", C.c_int), ("dE", C.c_int), ("out_slots", C.c_int), ("count", C.c_int),
                ("teams_per_group", C.c_int), ("group_threads", C.c_int), ("skew_cycles", C.c_int), ("evw", u32p), ("para", C.c_int)]


class MillerFixedArgs(C.Structure):
    _fields_ = [("lines", u32p), ("Ex", u32p), ("Ey", u32p), ("Einf", u8p), ("out_re", u32p), ("out_im", u32p),
                ("count", C.c_int)]


class PairDuoArgs(C.Structure):
    _fields_ = [("Mx", u32p), ("My", u32p), ("Minf", u8p), ("Ex", u32p), ("Ey", u32p), ("Einf", u8p),
                ("out_re", u32p), ("out_im", u32p), ("count", C.c_int)]


class EncArgs(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_int64)), ("r_be", u8p), ("rbytes", C.c_int), ("tabP", u32p), ("tabQ", u32p), ("wbitsQ", C.c_int),
                ("X", u32p), ("Y", u32p), ("Z", u32p), ("count", C.c_size_t), ("N", C.c_size_t),
                ("bx", u32p), ("by", u32p), ("binf", u8p), ("edw", C.c_int)]


class GtBlindArgs(C.Structure):
    _fields_ = [("re", u32p), ("im", u32p), ("r_be", u8p), ("rbytes", C.c_int), ("tabE", u32p), ("ore", u32p),
                ("oim", u32p), ("count", C.c_size_t)]


CONV_MAXW = 64


class PolyConvArgs(C.Structure):
    _fields_ = [("x", u32p), ("y", u32p), ("inf", u8p), ("d", C.c_int), ("nw", C.c_int), ("j_begin", C.c_int),
                ("j_count", C.c_int), ("top_bit", C.c_int), ("negate", C.c_int), ("w", C.c_uint64 * CONV_MAXW), ("whi", C.c_uint64 * CONV_MAXW),
                ("X", u32p), ("Y", u32p), ("Z", u32p), ("count", C.c_size_t)]


class DecLucasArgs(C.Structure):
    _fields_ = [("re", u32p), ("im", u32p), ("count", C.c_size_t), ("elems", u32p), ("slots", u32p),
                ("hmask", C.c_uint32), ("S", C.c_uint32), ("mmax", C.c_uint64), ("out", C.POINTER(C.c_int64)),
                ("status", u8p)]


class NormArgs(C.Structure):
    _fields_ = [("X", u32p), ("Y", u32p), ("Z", u32p), ("scratch", u32p), ("count", C.c_size_t), ("N", C.c_size_t),
                ("G", C.c_int), ("ox", u32p), ("oy", u32p), ("o_estride", C.c_size_t), ("o_lstride", C.c_size_t),
                ("inf", u8p)]


class G1AddArgs(C.Structure):
    _fields_ = [("x1", u32p), ("y1", u32p), ("inf1", u8p), ("x2", u32p), ("y2", u32p), ("inf2", u8p),
                ("N1", C.c_size_t), ("N2", C.c_size_t), ("bcast1", C.c_int), ("subtract", C.c_int),
                ("X", u32p), ("Y", u32p), ("Z", u32p), ("count", C.c_size_t), ("N", C.c_size_t)]


class G1AffAddArgs(C.Structure):
    _fields_ = [("x1", u32p), ("y1", u32p), ("inf1", u8p), ("x2", u32p), ("y2", u32p), ("inf2", u8p),
                ("bcast1", C.c_int), ("subtract", C.c_int), ("ox", u32p), ("oy", u32p), ("oinf", u8p),
                ("scratch", u32p), ("count", C.c_size_t), ("G", C.c_int)]


class G1MulArgs(C.Structure):
    _fields_ = [("x", u32p), ("y", u32p), ("inf", u8p), ("Nin", C.c_size_t), ("k_be", u8p), ("kbytes", C.c_int),
                ("X", u32p), ("Y", u32p), ("Z", u32p), ("count", C.c_size_t), ("N", C.c_size_t)]


class GtBinArgs(C.Structure):
    _fields_ = [("are", u32p), ("aim", u32p), ("bre", u32p), ("bim", u32p), ("Na", C.c_size_t), ("Nb", C.c_size_t),
                ("conj_b", C.c_int), ("ore", u32p), ("oim", u32p), ("count", C.c_size_t), ("N", C.c_size_t)]


class GtPowArgs(C.Structure):
    _fields_ = [("re", u32p), ("im", u32p), ("Nin", C.c_size_t), ("e_be", u8p), ("ebytes", C.c_int), ("mode", C.c_int),
                ("ore", u32p), ("oim", u32p), ("count", C.c_size_t), ("N", C.c_size_t)]


class BsgsBuildArgs(C.Structure):
    _fields_ = [("gen", u32p), ("elems", u32p), ("slots", u32p), ("hmask", C.c_uint32), ("S", C.c_uint32),
                ("chunk", C.c_int)]


class BsgsLookupArgs(C.Structure):
    _fields_ = [("re", u32p), ("im", u32p), ("Nin", C.c_size_t), ("count", C.c_size_t), ("elems", u32p),
                ("slots", u32p), ("hmask", C.c_uint32), ("S", C.c_uint32), ("ginv", u32p), ("giant_steps", C.c_uint32),
                ("mmax", C.c_uint64), ("out", C.POINTER(C.c_int64)), ("status", u8p)]


_lib = None


def build(loop: Optional[int] = None, extra: Sequence[str] = (), tag: str = "") -> str:
    """(Re)build libhostsim.so when a device header is newer than it.  `loop` builds a variant whose
    fused Miller routines use Fp::mul_loop<loop> (0 = unrolled products) instead of the default."""
    so = os.path.join(HERE, "libhostsim%s%s.so" % ("" if loop is None else "_u%d" % loop, tag))
    src = os.path.join(HERE, "hostsim.cpp")
    csrc = os.path.join(ROOT, "bgn_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        flags = [] if loop is None else ["-DBGN_MILLER_LOOP=%d" % loop, "-DBGN_MILLER_LOOP_A=%d" % loop]
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread"] + flags + list(extra) + ["-o", so, src])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def use_variant(loop: Optional[int], extra: Sequence[str] = (), tag: str = ""):
    """Switch the simulator to a build variant (loop shape and/or extra -D flags; None = default
    build); callers re-activate their Sim afterwards (constants live in the library)."""
    global _lib
    _lib = C.CDLL(build(loop, extra, tag))
    return _lib


def P32(a: np.ndarray):
    return a.ctypes.data_as(u32p)


def P8(a: np.ndarray):
    return a.ctypes.data_as(u8p)


def naf_digits(n: int) -> List[int]:
    d = []
    while n:
        z = 0
        if n & 1:
            z = 2 - (n & 3)
            n -= z
        d.append(z)
        n >>= 1
    return d[::-1]


def pick_L(p: int) -> int:
    for L in PRODUCT_L:
        if 32 * L >= p.bit_length() + 8:
            return L
    raise ValueError("p too large")


class Sim:
    """Orchestrates the simulated kernels the way api.cu does on the GPU."""

    def __init__(self, par, P, Q, q1: Optional[int] = None):
        self.par = par
        self.p = par.p
        self.L = pick_L(par.p)
        assert self.L in SUPPORTED_L, "hostsim only instantiates L in %r" % (SUPPORTED_L,)
        self.R = 1 << (32 * self.L)
        self.Rinv = pow(self.R, -1, self.p)
        self.B = par.coord_bytes
        self.nbytes = (par.n.bit_length() + 7) // 8
        fc = FieldConsts()
        for name, v in (("p", self.p), ("p2", 2 * self.p), ("one", self.R % self.p), ("r2", self.R * self.R % self.p),
                        ("p4", 4 * self.p), ("p8", 8 * self.p), ("p16", 16 * self.p)):
            arr = getattr(fc, name)
            for i in range(MAXL):
                arr[i] = (v >> (32 * i)) & 0xFFFFFFFF
        fc.np0 = (-pow(self.p, -1, 1 << 32)) % (1 << 32)
        pc = PairConsts()
        pc.l = par.l
        naf = naf_digits(par.n)
        pc.naf_len = len(naf)
        for i, z in enumerate(naf):
            pc.naf[i] = z
        if q1 is not None:
            pc.exp_bits = q1.bit_length()
            for i in range(MAX_EXPW):
                pc.exp[i] = (q1 >> (32 * i)) & 0xFFFFFFFF
            qn = naf_digits(q1)
            pc.exp_naf_len = len(qn)
            for i, z in enumerate(qn):
                pc.exp_naf[i] = z
        self.fc, self.pc = fc, pc
        self.P, self.Q = P, Q
        self.activate()

    def activate(self):
        lib().hs_set_consts(C.byref(self.fc), C.byref(self.pc), self.L)

    # ---- conversions between Python integers and Montgomery SoA
    def soa(self, vals: Sequence[int], mont: bool = True) -> np.ndarray:
        n = max(1, len(vals))
        a = np.zeros((n, self.L), dtype=np.uint32)  # array of elements: [N][L]
        for e, v in enumerate(vals):
            if mont:
                v = v * self.R % self.p
            for j in range(self.L):
                a[e, j] = (v >> (32 * j)) & 0xFFFFFFFF
        # range tracker: arrays built here hold canonical values (< p)
        lib().hs_track_array(P32(a), C.c_size_t(n), self.L, C.c_double(1.0))
        return a

    def range_report(self, reset: bool = True):
        """(largest bound attached in multiples of p, headroom R/p assumed, untracked reads, violations)"""
        out = (C.c_double * 4)()
        lib().hs_range_report(out, 1 if reset else 0)
        return out[0], out[1], int(out[2]), int(out[3])

    def unsoa(self, a: np.ndarray, count: int, mont: bool = True) -> List[int]:
        out = []
        for e in range(count):
            v = 0
            for j in range(self.L):
                v |= int(a[e, j]) << (32 * j)
            out.append(v * self.Rinv % self.p if mont else v)
        return out

    def g1_arrays(self, pts):
        x = self.soa([0 if P is None else P[0] for P in pts])
        y = self.soa([0 if P is None else P[1] for P in pts])
        inf = np.array([1 if P is None else 0 for P in pts] or [0], dtype=np.uint8)
        return x, y, inf

    # ---- kernels
    def miller(self, M, dM, E, dE, count, out_slots, e_bcast=False, teams_per_block=2, wide=False, para=None):
        Mx, My, Mi = self.g1_arrays(M)
        Ex, Ey, Ei = self.g1_arrays(E)
        nout = count * out_slots
        ore = np.zeros((nout, self.L), dtype=np.uint32)
        oim = np.zeros((nout, self.L), dtype=np.uint32)
        a = MillerArgs(P32(Mx), P32(My), P8(Mi), P32(Ex), P32(Ey), P8(Ei), None, P32(ore), P32(oim), Mx.shape[0],
                       Ex.shape[0], nout, 1 if e_bcast else 0, dM, dE, out_slots, count, teams_per_block,
                       teams_per_block * dE + 1, 0, None, (1 if dE >= 3 else 0) if para is None else para)  # (one idle thread per group)
        groups = 2 if count > teams_per_block else 1
        nt = groups * (teams_per_block * dE + 1)
        nblocks = (count + groups * teams_per_block - 1) // (groups * teams_per_block)
        assert (lib().hs_miller_wide if wide else lib().hs_miller)(self.L, C.byref(a), nblocks, nt) == 0
        return list(zip(self.unsoa(ore, nout), self.unsoa(oim, nout)))

    def miller_split(self, M, dM, E, dE, count, out_slots, teams_per_block=2, para=1):
        """k_miller_split (teamsplit.cuh): two threads per output-slot pair"""
        Mx, My, Mi = self.g1_arrays(M)
        Ex, Ey, Ei = self.g1_arrays(E)
        nout = count * out_slots
        ore = np.zeros((nout, self.L), dtype=np.uint32)
        oim = np.zeros((nout, self.L), dtype=np.uint32)
        nt = 2 * (teams_per_block * dE + 2)  # two halves (one role each) with idle threads: the inactive path
        a = MillerArgs(P32(Mx), P32(My), P8(Mi), P32(Ex), P32(Ey), P8(Ei), None, P32(ore), P32(oim), Mx.shape[0],
                       Ex.shape[0], nout, 0, dM, dE, out_slots, count, teams_per_block, nt, 0, None, para)
        nblocks = (count + teams_per_block - 1) // teams_per_block
        assert lib().hs_miller_split(self.L, C.byref(a), nblocks, nt) == 0
        return list(zip(self.unsoa(ore, nout), self.unsoa(oim, nout)))

    def multpoly_split(self, c1, d1, c2, d2, count, **kw):
        if d1 <= d2:
            return self.miller_split(c1, d1, c2, d2, count, d1 + d2, **kw)
        return self.miller_split(c2, d2, c1, d1, count, d1 + d2, **kw)

    def record_lines(self, base):
        """api.cu: ensure_linesP -- k_miller_record"""
        nsteps = lib().hs_miller_nsteps(self.L)
        ns = naf_digits(self.par.n)
        assert nsteps == sum(1 + (1 if (ns[i] != 0 and i != len(ns) - 1) else 0) for i in range(1, len(ns)))
        lines = np.zeros(nsteps * 2 * self.L, dtype=np.uint32)  # [step][cR / bI | aR / bI]
        scratch = np.zeros_like(lines)
        ok = C.c_int(0)
        bx, by = self.soa([base[0]]), self.soa([base[1]])
        assert lib().hs_miller_record(self.L, P32(bx), P32(by), P32(lines), P32(scratch), C.byref(ok)) == 0
        return lines if ok.value else None

    def pair_fixed(self, lines, Epts, nt=4):
        """k_miller_fixed: e(E[i], base) through the recorded line table"""
        count = len(Epts)
        Ex, Ey, Ei = self.g1_arrays(Epts)
        ore = np.zeros((count, self.L), dtype=np.uint32)
        oim = np.zeros_like(ore)
        a = MillerFixedArgs(P32(lines), P32(Ex), P32(Ey), P8(Ei), P32(ore), P32(oim), count)
        assert lib().hs_miller_fixed(self.L, C.byref(a), nt) == 0
        return list(zip(self.unsoa(ore, count), self.unsoa(oim, count)))

    def pair_fixed_pair(self, lines, Epts):
        """k_miller_fixed_pair: the same pairing, a lane pair per evaluation point (pairlane.cuh)"""
        count = len(Epts)
        Ex, Ey, Ei = self.g1_arrays(Epts)
        ore = np.zeros((count, self.L), dtype=np.uint32)
        oim = np.zeros_like(ore)
        a = MillerFixedArgs(P32(lines), P32(Ex), P32(Ey), P8(Ei), P32(ore), P32(oim), count)
        assert lib().hs_miller_fixed_pair(self.L, C.byref(a)) == 0
        return list(zip(self.unsoa(ore, count), self.unsoa(oim, count)))

    def pair_duo(self, A, Bp, np_=3):
        """k_pair_duo: e(A[i], B[i]) on two warps per group of pairings (pairwarp.cuh)"""
        count = len(A)
        Mx, My, Mi = self.g1_arrays(A)
        Ex, Ey, Ei = self.g1_arrays(Bp)
        ore = np.zeros((count, self.L), dtype=np.uint32)
        oim = np.zeros_like(ore)
        a = PairDuoArgs(P32(Mx), P32(My), P8(Mi), P32(Ex), P32(Ey), P8(Ei), P32(ore), P32(oim), count)
        assert lib().hs_pair_duo(self.L, C.byref(a), np_) == 0
        return list(zip(self.unsoa(ore, count), self.unsoa(oim, count)))

    def multpoly(self, c1, d1, c2, d2, count, wide=False):
        if d1 <= d2:
            return self.miller(c1, d1, c2, d2, count, d1 + d2, wide=wide)
        return self.miller(c2, d2, c1, d1, count, d1 + d2, wide=wide)

    def pair(self, A, Bp):
        return self.miller(A, 1, Bp, 1, len(A), 1, teams_per_block=3)

    def normalize(self, X, Y, Z, count, G=None):
        N = X.shape[0]
        scratch = np.zeros_like(X)
        ox = np.zeros_like(X)
        oy = np.zeros_like(X)
        inf = np.zeros(max(1, count), dtype=np.uint8)
        G = G or max(1, min(count, 3))
        a = NormArgs(P32(X), P32(Y), P32(Z), P32(scratch), count, N, G, P32(ox), P32(oy), self.L, 1, P8(inf))
        assert lib().hs_normalize(self.L, C.byref(a)) == 0
        xs, ys = self.unsoa(ox, count), self.unsoa(oy, count)
        return [None if inf[e] else (xs[e], ys[e]) for e in range(count)]

    def build_table(self, base, nwin, hb=8):
        """k_tab_bases + k_tab_fill + k_normalize (api.cu: build_table_with): nwin windows of hb bits"""
        L = self.L
        bx, by = self.soa([base[0]]), self.soa([base[1]])
        X = np.zeros((nwin, L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        assert lib().hs_tab_bases(L, P32(bx), P32(by), nwin, hb, P32(X), P32(Y), P32(Z), nwin) == 0
        bases = self.normalize(X, Y, Z, nwin)
        ax, ay, ainf = self.g1_arrays(bases)
        nent = nwin * ((1 << hb) - 1)
        X = np.zeros((nent, L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        assert lib().hs_tab_fill(L, P32(ax), P32(ay), P8(ainf), nwin, nwin, hb, P32(X), P32(Y), P32(Z), nent) == 0
        tab = np.zeros(nent * 2 * L, dtype=np.uint32)
        scratch = np.zeros_like(X)
        a = NormArgs(P32(X), P32(Y), P32(Z), P32(scratch), nent, nent, 7, P32(tab), P32(tab[L:]), 2 * L, 1, None)
        assert lib().hs_normalize(L, C.byref(a)) == 0
        return tab

    def build_table16(self, tab8, nwin8):
        """16-bit windows from the 8-bit table of the key"""
        return self.build_table_wide(tab8, nwin8, (nwin8 + 1) // 2, 2, 8)

    def build_table_wide(self, tabh, nwin_h, nwin, nsub, hb):
        """k_tabw_fill + k_normalize: nwin windows of nsub * hb bits from a table of nwin_h windows of hb bits
        (api.cu: ensure_tabQw), built in two chunks the way the large tables are"""
        L = self.L
        nent = nwin * ((1 << (nsub * hb)) - 1)
        X = np.zeros((nent, L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        half = nent // 2 + 3
        assert lib().hs_tabw_fill(L, P32(tabh), nwin_h, nsub, hb, P32(X), P32(Y), P32(Z), C.c_size_t(0), C.c_size_t(half)) == 0
        assert lib().hs_tabw_fill(L, P32(tabh), nwin_h, nsub, hb, P32(X[half:]), P32(Y[half:]), P32(Z[half:]),
                                  C.c_size_t(half), C.c_size_t(nent - half)) == 0
        tab = np.zeros(nent * 2 * L, dtype=np.uint32)
        scratch = np.zeros_like(X)
        a = NormArgs(P32(X), P32(Y), P32(Z), P32(scratch), nent, nent, 64, P32(tab), P32(tab[L:]), 2 * L, 1, None)
        assert lib().hs_normalize(L, C.byref(a)) == 0
        return tab

    def table_to_edwards(self, tabw, G=5):
        """k_tab_edwards (api.cu: table_to_edwards): x || y entries -> u || v || u v entries, or None if an
        entry has no Edwards image"""
        L = self.L
        count = len(tabw) // (2 * L)
        tabe = np.zeros(count * 3 * L, dtype=np.uint32)
        scratch = np.zeros(count * L, dtype=np.uint32)
        bad = C.c_int(0)
        assert lib().hs_tab_edwards(L, P32(tabw), P32(tabe), P32(scratch), C.c_size_t(count), G, C.byref(bad)) == 0
        return None if bad.value else tabe

    def encrypt(self, xs, rs, tabP, tabQ, wbitsQ=8, base=None, edw=False):
        """base: optional list of starting points (bgn_g1_blind_batch: base + r*Q; xs then None);
        edw: the tables hold twisted Edwards points (table_to_edwards)"""
        count = len(rs) if xs is None else len(xs)
        x = np.array(xs if xs is not None else [0], dtype=np.int64)
        bx = by = binf = None
        if base is not None:
            bxa, bya, binfa = self.g1_arrays(base)
            bx, by, binf = P32(bxa), P32(bya), P8(binfa)
        X = np.zeros((max(1, count), self.L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        if rs is None:
            rbuf, rp = None, None
        else:
            rbuf = np.frombuffer(b"".join(int(r).to_bytes(self.nbytes, "big") for r in rs), dtype=np.uint8).copy()
            rp = P8(rbuf)
        a = EncArgs(x.ctypes.data_as(C.POINTER(C.c_int64)) if xs is not None else None, rp, self.nbytes, P32(tabP),
                    P32(tabQ), wbitsQ, P32(X), P32(Y), P32(Z), count, count, bx, by, binf, 1 if edw else 0)
        assert lib().hs_encrypt(self.L, C.byref(a)) == 0
        return self.normalize(X, Y, Z, count)

    def g1_add(self, A, Bp, subtract=False, bcast1=False):
        count = len(Bp)
        x1, y1, i1 = self.g1_arrays(A)
        x2, y2, i2 = self.g1_arrays(Bp)
        X = np.zeros((max(1, count), self.L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        a = G1AddArgs(P32(x1), P32(y1), P8(i1), P32(x2), P32(y2), P8(i2), x1.shape[0], x2.shape[0],
                      1 if bcast1 else 0, 1 if subtract else 0, P32(X), P32(Y), P32(Z), count, count)
        assert lib().hs_g1_add(self.L, C.byref(a)) == 0
        return self.normalize(X, Y, Z, count)

    def g1_affadd(self, A, Bp, subtract=False, bcast1=False, G=3):
        """k_g1_affadd: EAdd / ESub in affine coordinates with one shared inversion per thread"""
        count = len(Bp)
        x1, y1, i1 = self.g1_arrays(A)
        x2, y2, i2 = self.g1_arrays(Bp)
        ox = np.zeros((max(1, count), self.L), dtype=np.uint32)
        oy = np.zeros_like(ox)
        oinf = np.zeros(max(1, count), dtype=np.uint8)
        scratch = np.zeros_like(ox)
        a = G1AffAddArgs(P32(x1), P32(y1), P8(i1), P32(x2), P32(y2), P8(i2), 1 if bcast1 else 0, 1 if subtract else 0,
                         P32(ox), P32(oy), P8(oinf), P32(scratch), count, max(1, min(G, count)))
        assert lib().hs_g1_affadd(self.L, C.byref(a)) == 0
        xs, ys = self.unsoa(ox, count), self.unsoa(oy, count)
        return [None if oinf[e] else (xs[e], ys[e]) for e in range(count)]

    def g1_mulvar(self, A, ks, kbytes):
        count = len(A)
        x, y, inf = self.g1_arrays(A)
        kb = np.frombuffer(b"".join(int(k).to_bytes(kbytes, "big") for k in ks), dtype=np.uint8).copy()
        X = np.zeros((max(1, count), self.L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        a = G1MulArgs(P32(x), P32(y), P8(inf), x.shape[0], P8(kb), kbytes, P32(X), P32(Y), P32(Z), count, count)
        assert lib().hs_g1_mulvar(self.L, C.byref(a)) == 0
        return self.normalize(X, Y, Z, count)

    def gt_arrays(self, vals):
        return self.soa([v[0] for v in vals]), self.soa([v[1] for v in vals])

    def gt_mul(self, A, Bv, conj_b=False):
        count = len(A)
        are, aim = self.gt_arrays(A)
        bre, bim = self.gt_arrays(Bv)
        ore, oim = np.zeros_like(are), np.zeros_like(are)
        a = GtBinArgs(P32(are), P32(aim), P32(bre), P32(bim), count, count, 1 if conj_b else 0, P32(ore), P32(oim),
                      count, count)
        assert lib().hs_gt_mul(self.L, C.byref(a)) == 0
        return list(zip(self.unsoa(ore, count), self.unsoa(oim, count)))

    def gt_pow(self, A, mode, exps=None, ebytes=0, pair=False):
        """pair=True: k_gt_pow_pair (fixed exponent, a lane pair per element)"""
        count = len(A)
        are, aim = self.gt_arrays(A)
        ore, oim = np.zeros_like(are), np.zeros_like(are)
        eb, ep = None, None
        if exps is not None:
            eb = np.frombuffer(b"".join(int(k).to_bytes(ebytes, "big") for k in exps), dtype=np.uint8).copy()
            ep = P8(eb)
        a = GtPowArgs(P32(are), P32(aim), count, ep, ebytes, mode, P32(ore), P32(oim), count, count)
        if pair:
            lib().hs_track_array(P32(are), C.c_size_t(count), self.L, C.c_double(2.0))
            lib().hs_track_array(P32(aim), C.c_size_t(count), self.L, C.c_double(2.0))
            assert mode == 1 and lib().hs_gt_pow_pair(self.L, C.byref(a)) == 0
        else:
            assert lib().hs_gt_pow(self.L, C.byref(a)) == 0
        return list(zip(self.unsoa(ore, count), self.unsoa(oim, count)))

    def gt_reduce(self, vals, nterms, ncoeff, G):
        re, im = self.gt_arrays(vals)
        ore = np.zeros((G * ncoeff, self.L), dtype=np.uint32)
        oim = np.zeros_like(ore)
        assert lib().hs_gt_reduce(self.L, P32(re), P32(im), re.shape[0], nterms, ncoeff, G, P32(ore), P32(oim),
                                  G * ncoeff) == 0
        return list(zip(self.unsoa(ore, G * ncoeff), self.unsoa(oim, G * ncoeff)))

    def gt_table(self, gen: Tuple[int, int], nwin: int) -> np.ndarray:
        """k_gt_tab_bases + k_gt_tab_fill (api.cu: ensure_tabE)"""
        L = self.L
        g = np.concatenate([self.soa([gen[0]])[0], self.soa([gen[1]])[0]]).astype(np.uint32)
        lib().hs_track_array(P32(g), C.c_size_t(2), L, C.c_double(1.0))
        bases = np.zeros(nwin * 2 * L, dtype=np.uint32)
        tab = np.zeros(nwin * 255 * 2 * L, dtype=np.uint32)
        assert lib().hs_gt_tab_bases(L, P32(g), nwin, P32(bases)) == 0
        assert lib().hs_gt_tab_fill(L, P32(bases), nwin, P32(tab)) == 0
        return tab

    def gt_blind(self, A, rs, tabE):
        count = len(A)
        are, aim = self.gt_arrays(A)
        ore, oim = np.zeros_like(are), np.zeros_like(are)
        rbuf = np.frombuffer(b"".join(int(r).to_bytes(self.nbytes, "big") for r in rs), dtype=np.uint8).copy()
        a = GtBlindArgs(P32(are), P32(aim), P8(rbuf), self.nbytes, P32(tabE), P32(ore), P32(oim), count)
        assert lib().hs_gt_blind(self.L, C.byref(a)) == 0
        return list(zip(self.unsoa(ore, count), self.unsoa(oim, count)))

    def polyconv(self, elems, d, is_l2, w, j_begin, j_count, negate, count):
        """k_g1_polyconv (+ k_normalize) / k_gt_polyconv as api.cu: polyconv() drives them"""
        nout = count * j_count
        a = PolyConvArgs()
        a.d, a.nw, a.j_begin, a.j_count, a.negate, a.count = d, len(w), j_begin, j_count, 1 if negate else 0, count
        anyw = 0
        for k, v in enumerate(w):
            a.w[k] = v & 0xFFFFFFFFFFFFFFFF
            a.whi[k] = v >> 64
            anyw |= v
        a.top_bit = anyw.bit_length() - 1
        X = np.zeros((max(1, nout), self.L), dtype=np.uint32)
        Y = np.zeros_like(X)
        Z = np.zeros_like(X)
        a.X, a.Y, a.Z = P32(X), P32(Y), P32(Z)
        if is_l2:
            re, im = self.gt_arrays(elems)
            a.x, a.y = P32(re), P32(im)
            assert lib().hs_gt_polyconv(self.L, C.byref(a)) == 0
            return list(zip(self.unsoa(X, nout), self.unsoa(Y, nout)))
        x, y, inf = self.g1_arrays(elems)
        a.x, a.y, a.inf = P32(x), P32(y), P8(inf)
        assert lib().hs_g1_polyconv(self.L, C.byref(a)) == 0
        return self.normalize(X, Y, Z, nout)

    def gt_to_bytes_padded(self, vals, grp: int, pad: int) -> bytes:
        re, im = self.gt_arrays(vals)
        count = (len(vals) // grp) * (grp + pad)
        out = np.zeros(count * 2 * self.B, dtype=np.uint8)
        assert lib().hs_fp2_to_bytes(self.L, P32(re), P32(im), re.shape[0], count, P8(out), self.B, grp, pad) == 0
        return out.tobytes()

    def bsgs_setup(self, gsk: Tuple[int, int], msg_space: int, S: Optional[int] = None):
        import math
        L, p = self.L, self.p
        bound = int(math.ceil(math.sqrt(float(msg_space))))
        self.mmax = bound * bound + bound + 2
        S = S or self.mmax
        hs = 1
        while hs < 2 * S:
            hs <<= 1
        gen = np.concatenate([self.soa([gsk[0]])[0], self.soa([gsk[1]])[0]]).astype(np.uint32)
        self.bs_elems = np.zeros(S * 2 * L, dtype=np.uint32)
        self.bs_slots = np.zeros(hs, dtype=np.uint32)
        a = BsgsBuildArgs(P32(gen), P32(self.bs_elems), P32(self.bs_slots), hs - 1, S, 7)
        assert lib().hs_bsgs_build(L, C.byref(a)) == 0
        from oracle import bgn_oracle as O
        gi = O.fp2_conj(O.fp2_pow(gsk, S, p), p)
        self.bs_ginv = np.concatenate([self.soa([gi[0]])[0], self.soa([gi[1]])[0]]).astype(np.uint32)
        self.bs_S, self.bs_hmask = S, hs - 1
        self.bs_giant = (self.mmax + S - 1) // S

    def bsgs_lookup(self, csk):
        count = len(csk)
        re, im = self.gt_arrays(csk)
        out = np.zeros(count, dtype=np.int64)
        status = np.zeros(count, dtype=np.uint8)
        a = BsgsLookupArgs(P32(re), P32(im), count, count, P32(self.bs_elems), P32(self.bs_slots), self.bs_hmask,
                           self.bs_S, P32(self.bs_ginv), self.bs_giant, self.mmax,
                           out.ctypes.data_as(C.POINTER(C.c_int64)), P8(status))
        assert lib().hs_bsgs_lookup(self.L, C.byref(a)) == 0
        return list(out), list(status)

    def dec_lucas(self, cts):
        """k_dec_lucas: Lucas ladder for C^q1 + search by real part (needs bsgs_setup with S = mmax)"""
        assert self.bs_giant == 1
        count = len(cts)
        re, im = self.gt_arrays(cts)
        lib().hs_track_array(P32(re), C.c_size_t(count), self.L, C.c_double(2.0))
        lib().hs_track_array(P32(im), C.c_size_t(count), self.L, C.c_double(2.0))
        out = np.zeros(count, dtype=np.int64)
        status = np.zeros(count, dtype=np.uint8)
        a = DecLucasArgs(P32(re), P32(im), count, P32(self.bs_elems), P32(self.bs_slots), self.bs_hmask, self.bs_S,
                         self.mmax, out.ctypes.data_as(C.POINTER(C.c_int64)), P8(status))
        assert lib().hs_dec_lucas(self.L, C.byref(a)) == 0
        return list(out), list(status)

    # ---- byte formats
    def g1_from_bytes(self, data: bytes, count: int):
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        x = np.zeros((count, self.L), dtype=np.uint32)
        y = np.zeros_like(x)
        inf = np.zeros(count, dtype=np.uint8)
        assert lib().hs_g1_from_bytes(self.L, P8(buf), self.B, count, P32(x), P32(y), P8(inf), count) == 0
        xs, ys = self.unsoa(x, count), self.unsoa(y, count)
        return [None if inf[e] else (xs[e], ys[e]) for e in range(count)]

    def g1_to_bytes(self, pts) -> bytes:
        count = len(pts)
        x, y, inf = self.g1_arrays(pts)
        out = np.zeros(count * 2 * self.B, dtype=np.uint8)
        assert lib().hs_g1_to_bytes(self.L, P32(x), P32(y), P8(inf), x.shape[0], count, P8(out), self.B) == 0
        return out.tobytes()

    def gt_roundtrip_bytes(self, vals) -> bytes:
        count = len(vals)
        re, im = self.gt_arrays(vals)
        out = np.zeros(count * 2 * self.B, dtype=np.uint8)
        assert lib().hs_fp2_to_bytes(self.L, P32(re), P32(im), count, count, P8(out), self.B, 0, 0) == 0
        re2, im2 = np.zeros_like(re), np.zeros_like(im)
        assert lib().hs_fp2_from_bytes(self.L, P8(out), self.B, count, P32(re2), P32(im2), count) == 0
        back = list(zip(self.unsoa(re2, count), self.unsoa(im2, count)))
        assert back == [(v[0] % self.p, v[1] % self.p) for v in vals]
        return out.tobytes()
