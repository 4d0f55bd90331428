"""NewKeyGen (bgn.go:65-138) for the host mirror.

The reference generates keys with crypto/rand + PBC (pbc.GenerateA1, G1.Rand, PowBig).  Here the
one-off host part -- two random primes, the type-A1 parameters (the smallest l = 0 mod 4 with
p = l*n - 1 prime), one square root mod p for a random curve point -- is plain Python integer
arithmetic, and every group operation (the cofactor and generator scalar multiplications, the
generator test of findGenerator) runs on the GPU through the C-ABI like everything else.  In a Go
integration key generation stays in Go; this exists so that the Python mirror is usable on its own.
"""
from __future__ import annotations

import secrets
from typing import Callable, Optional, Tuple

import numpy as np

_SMALL_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97)


class _SystemRng:
    """crypto/rand stand-in with the two calls the generator needs"""

    def getrandbits(self, k: int) -> int:
        return secrets.randbits(k)

    def randrange(self, n: int) -> int:
        return secrets.randbelow(n)


def is_probable_prime(m: int, rng, rounds: int = 24) -> bool:
    """trial division + Miller-Rabin (math/big.ProbablyPrime(20) in the reference's rand.Prime)"""
    if m < 2:
        return False
    for q in _SMALL_PRIMES:
        if m % q == 0:
            return m == q
    d, s = m - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for _ in range(rounds):
        a = 2 + rng.randrange(m - 3)
        x = pow(a, d, m)
        if x in (1, m - 1):
            continue
        for _ in range(s - 1):
            x = x * x % m
            if x == m - 1:
                break
        else:
            return False
    return True


def rand_prime(rng, bits: int) -> int:
    """crypto/rand.Prime: `bits` bits with the top two set (so a product of two has exactly 2*bits bits)"""
    while True:
        c = rng.getrandbits(bits) | (3 << (bits - 2)) | 1
        if is_probable_prime(c, rng):
            return c


def a1_params(n: int, rng) -> Tuple[int, int]:
    """pbc_param_init_a1_gen: the smallest multiple of 4, l, with p = l*n - 1 prime -> (p, l)"""
    l = 4
    while not is_probable_prime(l * n - 1, rng):
        l += 4
    return l * n - 1, l


def NewKeyGen(keyBits: int, msgSpace: int, polyBase: int = 3, fpScaleBase: int = 3, fpPrecision: float = 0.0001,
              deterministic: bool = True, rng=None, device: int = 0, engine_factory: Optional[Callable] = None):
    """-> (PublicKey, SecretKey).  `rng` (getrandbits / randrange) injects the randomness; the default
    draws from the operating system like crypto/rand.  Raises ValueError where the reference panics."""
    from .bgn import PublicKey, SecretKey
    from .engine import Engine
    if keyBits < 16:
        raise ValueError("key bits must be >= 16 bits in length")  # bgn.go:67-69
    if keyBits % 2 != 0:
        raise ValueError("key bits must be divisible by 2")  # bgn.go:71-73
    rng = rng or _SystemRng()
    make = engine_factory or (lambda p, n, l, P, Q: Engine(p, n, l, P, Q, device))
    q1 = rand_prime(rng, keyBits // 2)  # newPrimeTuple, bgn.go:151-168
    q2 = rand_prime(rng, keyBits // 2)
    while q2 == q1:
        q2 = rand_prime(rng, keyBits // 2)
    if q1 < msgSpace or q2 < msgSpace:
        raise ValueError("Message space is greater than the group order!")  # bgn.go:87-89
    n = q1 * q2
    p, l = a1_params(n, rng)  # pbc.GenerateA1, bgn.go:93
    B = (p.bit_length() + 7) // 8
    nb = (n.bit_length() + 7) // 8

    def scal(k: int, width: int) -> np.ndarray:
        return np.frombuffer(k.to_bytes(width, "big"), dtype=np.uint8)

    while True:
        # G1.Rand (pbc curve_random): a random point of the curve y^2 = x^3 + x, times the cofactor l
        x = rng.randrange(p)
        rhs = (x * x * x + x) % p
        y = pow(rhs, (p + 1) // 4, p)  # p = 3 (mod 4)
        if y * y % p != rhs or (x == 0 and y == 0):
            continue
        if rng.getrandbits(1):
            y = (-y) % p
        raw = x.to_bytes(B, "big") + y.to_bytes(B, "big")
        boot = make(p, n, l, raw, raw)  # any curve point serves as the bootstrap context's generators
        try:
            pt = np.frombuffer(raw, dtype=np.uint8)
            G = boot.g1_mulconst_batch(pt, scal(l, 8), 8)
            # findGenerator (bgn.go:170-192): q1*G != O and n*G == O
            t1 = boot.g1_mulconst_batch(G, scal(q1, nb), nb)
            t2 = boot.g1_mulconst_batch(G, scal(n, nb), nb)
            if not t1.any() or t2.any():
                continue
            P = boot.g1_mulconst_batch(G, scal(4 * l, 9), 9)  # bgn.go:113
            R = rng.randrange(n)  # bgn.go:117
            Q = boot.g1_mulconst_batch(boot.g1_mulconst_batch(P, scal(R, nb), nb), scal(q2, nb), nb)  # bgn.go:118-119
            P, Q = bytes(P.tobytes()), bytes(Q.tobytes())
        finally:
            boot.close()
        break
    pk = PublicKey(p, n, l, P, Q, msgSpace, Deterministic=deterministic, polyBase=polyBase, fpScaleBase=fpScaleBase,
                   fpPrecision=fpPrecision, device=device, engine=make(p, n, l, P, Q))
    return pk, SecretKey(q1, R, polyBase)
