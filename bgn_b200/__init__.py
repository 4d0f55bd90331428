"""bgn_b200 -- B200-native batched engine for the BGN cryptosystem.

`bgn_b200.bgn` mirrors the reference's Go API (sachaservan/bgn); every group
operation runs in hand-written sm_100a CUDA behind the C-ABI of
include/bgn_b200.h (bgn_b200/libbgn_b200.so).  No CPU fallback exists: using a
compute entry point without the built CUDA library raises.
"""
from .engine import BgnError, DeviceBatch, Engine, EngineGroup, bench_imad_peak  # noqa: F401
from .bgn import (Ciphertext, DLError, PolyCiphertext, PolyCiphertextBatch, PublicKey, SecretKey)  # noqa: F401
from .gadgets import DecryptionProof, NewDecryptionProof, ProofOfPlaintextKnowledge  # noqa: F401
from .keygen import NewKeyGen  # noqa: F401
from .plaintext import PolyEncodingParams, PolyPlaintext  # noqa: F401
