"""The C port of the oracle (oracle/cpu_ref.c: what bench.py times as the CPU baseline) must
byte-match the golden vectors and the Python oracle.  No GPU."""
import random

import numpy as np
import pytest

from conftest import load_golden, unhex
from oracle import bgn_oracle as O
from oracle.cpu_ref import CpuRef


def buf(xs):
    return np.frombuffer(unhex(xs), dtype=np.uint8)


def test_cpu_ref_golden(golden):
    g = golden
    R = CpuRef(int(g["p"], 16), int(g["n"], 16), g["l"], threads=3)
    assert R.B == g["coord_bytes"]
    v = g["pair"]
    assert R.pair_batch(buf(v["a"]), buf(v["b"])).tobytes() == unhex(v["out"])
    v = g["multpoly"]
    assert R.multpoly_batch(buf(v["c1"]), v["d1"], buf(v["c2"]), v["d2"], 1).tobytes() == unhex(v["out"])
    v = g["encrypt"]
    r = np.frombuffer(b"".join(int(x, 16).to_bytes(R.nbytes, "big") for x in v["r"]), dtype=np.uint8)
    got = R.encrypt_batch(bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), v["x"], r)
    assert got.tobytes() == unhex(v["out"])
    v = g["decrypt_l2"]
    assert R.gt_pow_batch(buf(v["in"]), int(g["q1"], 16)).tobytes() == unhex(v["csk"])


@pytest.mark.parametrize("kb", [128, 512])
def test_cpu_ref_random_vs_python_oracle(kb):
    g = load_golden(kb)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
    R = CpuRef(par.p, par.n, par.l, threads=2)
    rng = random.Random(kb)
    cnt = 10 if kb == 128 else 3
    A = [O.g1_mul(rng.randrange(par.n), P, par.p) for _ in range(cnt)]
    Bp = [O.g1_mul(rng.randrange(par.n), P, par.p) for _ in range(cnt)]
    a = np.frombuffer(b"".join(O.g1_to_bytes(x, par) for x in A), dtype=np.uint8)
    b = np.frombuffer(b"".join(O.g1_to_bytes(x, par) for x in Bp), dtype=np.uint8)
    exp = b"".join(O.gt_to_bytes(O.pairing(x, y, par), par) for x, y in zip(A, Bp))
    assert R.pair_batch(a, b).tobytes() == exp
