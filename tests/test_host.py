"""CPU tests of the host side: the plaintext-encoding mirror (bgn_b200/plaintext.py vs the
oracle's restatement of plaintext.go), the C-ABI surface (library loads, exports every symbol
include/bgn_b200.h declares -- no compute calls here), and the "no CPU fallback" rule."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, load_golden
from oracle import bgn_oracle as O


def test_plaintext_mirror_matches_oracle():
    from bgn_b200 import plaintext as PT
    tab = PT.EncodingTable(3)
    deg, sums = O.compute_encoding_table(3)
    assert tab.degree == deg and tab.degree_sum == sums
    for v in list(range(0, 1500)) + [65535, 2 ** 31 - 1, 3 ** 40 + 17]:
        assert PT.balancedEncode(v, tab) == O.balanced_encode(v, 3)
        assert PT.unbalancedEncode(v, tab) == O.unbalanced_encode(v, 3)
        assert PT.balancedEncode(-v, tab) == O.balanced_encode(-v, 3)
    for x in (0.5, 0.123, 1 / 3, 0.9999, 0.2, 0.13, 0.12):
        assert PT.rationalize(x, 3, 0.0001) == O.rationalize(x, 3, 0.0001)
    prm = PT.PolyEncodingParams(3, 3, 0.0001)
    pk = O.PublicKey(O.A1Params(7, 2, 4), None, None, 1021)
    for m in (0.0, 1.0, 9.123, 100.1, 4.2, 0.1, 50.1, 41.2, 9.13, 4.12, 1.1, 40.2, 65535.0):
        a, b = PT.NewPolyPlaintext(m, prm, tab), pk.new_poly_plaintext(m)
        assert (a.Coefficients, a.Degree, a.ScaleFactor) == (b.coefficients, b.degree, b.scale_factor)
        assert a.PolyEval() == b.poly_eval()
        a, b = PT.NewUnbalancedPlaintext(m, prm, tab), pk.new_unbalanced_plaintext(m)
        assert (a.Coefficients, a.Degree, a.ScaleFactor) == (b.coefficients, b.degree, b.scale_factor)
    with pytest.raises(ValueError):
        PT.NewPolyPlaintext(-1.0, prm, tab)
    # the reference benchmark plaintext: 100.1 -> 13 slots -> 169 pairings per MultPoly
    assert PT.NewPolyPlaintext(100.1, prm, tab).Degree == 13


def header_symbols():
    with open(os.path.join(ROOT, "include", "bgn_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(bgn_[a-z0-9_]+)\s*\(", src)))


def test_cabi_exports_every_declared_symbol():
    from bgn_b200 import _cabi
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(_cabi.SIGNATURES) == syms, "bgn_b200/_cabi.py and include/bgn_b200.h disagree"
    lib = _cabi.load()  # dlopen works without a GPU (no CUDA call is made)
    for s in syms:
        assert getattr(lib, s) is not None


def test_cabi_argument_errors_without_gpu():
    """NULL-context / NULL-argument handling happens before any CUDA call."""
    from bgn_b200 import _cabi
    lib = _cabi.load()
    assert lib.bgn_ctx_info(None, None, None, None) == _cabi.BGN_E_BADARG
    assert lib.bgn_encrypt_batch(None, None, None, 1, None) == _cabi.BGN_E_BADARG
    assert lib.bgn_ctx_create(None, 0, ctypes.byref(ctypes.c_void_p())) == _cabi.BGN_E_BADARG
    # failures without a context are reported through the per-thread string (no stderr, no abort)
    assert lib.bgn_global_last_error() == b"bgn_ctx_create: null argument or l == 0"
    assert lib.bgn_last_error(None) == lib.bgn_global_last_error()
    assert lib.bgn_encrypt_batch(None, None, None, 1, None) == _cabi.BGN_E_BADARG
    assert lib.bgn_global_last_error() == b"null context"
    lib.bgn_ctx_destroy(None)  # no-op


def test_product_never_touches_the_oracle_and_has_no_cpu_fallback():
    pkg = os.path.join(ROOT, "bgn_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn + " imports the oracle"
                assert "hostsim.cpp" not in src or fn.endswith(".cuh")


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from bgn_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _cabi.load()


def test_work_model_numbers():
    from bgn_b200 import workmodel as W
    from conftest import load_golden
    g = load_golden(512)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    assert W.pick_limbs(p) == 17 and W.products_per_modmul(17) == 595
    unit = W.miller_unit_modmuls(p, n, l, 11, 11)
    canon = 121 * W.canonical_pairing_modmuls(n, l)
    assert unit < canon / 3  # sharing (lines, squarings, final exp) removes > 2/3 of the canonical work
    assert W.canonical_pairing_modmuls(n, l) == O.canonical_modmuls_per_pairing(O.A1Params(p, n, l))
    assert sum(d * 2 ** i for i, d in enumerate(reversed(W.naf_digits(n)))) == n


def test_element_string_format():
    """pbc's base-10 element format as printed by Ciphertext.String (ciphertext.go:60-72)."""
    from bgn_b200.bgn import Ciphertext, PolyCiphertext
    from oracle import bgn_oracle as O
    g = load_golden(64)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    raw = bytes.fromhex(g["encrypt"]["out"][2])
    pt = O.g1_from_bytes(raw, par)
    assert Ciphertext(raw, False).String() == O.g1_string(pt) + "\n"
    assert Ciphertext(bytes(len(raw)), False).String() == "O\n"
    gt = bytes.fromhex(g["pair"]["out"][0])
    assert Ciphertext(gt, True).String() == O.gt_string(O.gt_from_bytes(gt, par)) + "\n"
    pc = PolyCiphertext([Ciphertext(raw, False), Ciphertext(bytes(len(raw)), False)], 2, 0, False)
    assert pc.String() == O.g1_string(pt) + "\nO\n"
