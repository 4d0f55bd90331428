"""Zero-knowledge gadgets of the reference (gadgets.go) on top of the batched engine: a decryption
proof (the opening (v, r) of a level-1 ciphertext) and a Schnorr-style proof of plaintext knowledge.
Every group operation is a C-ABI call; the checks take LISTS so that a batch of proofs costs one
crossing per operation instead of one per proof (a further consumer of the
fixed-base kernels)."""
from __future__ import annotations

import hashlib
import secrets
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class DecryptionProof:  # gadgets.go:16-21
    Value: int
    Randomness: int


def NewDecryptionProof(v: int, r: int) -> DecryptionProof:  # gadgets.go:24-28
    return DecryptionProof(v, r)


@dataclass
class ProofOfPlaintextKnowledge:  # gadgets.go:10-14
    Ct: object  # Ciphertext
    Nonce: object  # Ciphertext
    DL: Optional[int]


def hash_proof(proof: ProofOfPlaintextKnowledge) -> int:
    """gadgets.go:80-96: SHA-256 over Ct.C.Bytes() || Nonce.C.Bytes(), read as a big-endian integer."""
    return int.from_bytes(hashlib.sha256(bytes(proof.Ct.C) + bytes(proof.Nonce.C)).digest(), "big")


def new_proof_of_plaintext_knowledge(pk, sk, v: int, z: int, nonce1: Optional[int] = None) -> ProofOfPlaintextKnowledge:
    """gadgets.go:32-56.  nonce1 injects the prover's randomness (newCryptoRandom(pk.N) in the reference)."""
    if nonce1 is None:
        nonce1 = secrets.randbelow(pk.N)
    ct = pk.EncryptWithRandomness(v, z)        # g^v h^z
    nonce = pk.EncryptWithRandomness(nonce1, 0)  # g^r
    proof = ProofOfPlaintextKnowledge(ct, nonce, None)
    c = hash_proof(proof)
    dl = nonce1 + c * v + sk.R * z * c * (pk.N // sk.Key)  # r + c v + R z c q2
    proof.DL = dl % pk.N
    return proof


def check_decryption_proofs(pk, cts: Sequence, proofs: Sequence[DecryptionProof]) -> List[bool]:
    """gadgets.go:59-63 over a batch: ct == EncryptWithRandomness(Value, Randomness)."""
    if not cts:
        return []
    res = pk._encrypt_big([p.Value for p in proofs], [p.Randomness for p in proofs])
    eb = pk.elem_bytes
    return [bytes(c.C) == res[i * eb:(i + 1) * eb] and not c.L2 for i, c in enumerate(cts)]


def check_proofs_of_plaintext_knowledge(pk, cts: Sequence, proofs: Sequence[ProofOfPlaintextKnowledge]) -> List[bool]:
    """gadgets.go:67-78 over a batch: ct^c * Nonce == P^DL with c = hash(proof)."""
    if not cts:
        return []
    eng, eb, n = pk.engine, pk.elem_bytes, len(cts)
    cs = [hash_proof(p) for p in proofs]
    cbuf = np.frombuffer(b"".join(bytes(c.C) for c in cts), dtype=np.uint8)
    lhs = eng.g1_mulconst_batch(cbuf, eng.scalars_be(cs, 32), 32)          # ct^c (c is a SHA-256 value)
    lhs = eng.g1_add_batch(lhs, np.frombuffer(b"".join(bytes(p.Nonce.C) for p in proofs), dtype=np.uint8))
    P = np.frombuffer(bytes(pk.P) * n, dtype=np.uint8)
    nb = eng.scalar_bytes
    rhs = eng.g1_mulconst_batch(P, eng.scalars_be([p.DL % pk.N for p in proofs], nb), nb)  # P^DL
    lhs, rhs = bytes(np.asarray(lhs).tobytes()), bytes(np.asarray(rhs).tobytes())
    return [lhs[i * eb:(i + 1) * eb] == rhs[i * eb:(i + 1) * eb] for i in range(n)]
