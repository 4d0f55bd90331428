"""encoding/gob envelopes (bgn_b200/gobwire.py) -- pinned by the worked example of the gob
documentation; the bgn wrappers round-trip and survive foreign type ids and field orders."""
import pytest

from bgn_b200 import gobwire as G

# "struct { A, B int }" named Point, value {22, 33}: the byte stream of the encoding/gob package documentation
DOC_EXAMPLE = bytes.fromhex(
    "1fff810301010550 6f696e7401ff8200 0102010158010400 0101590104000000 07ff82012c014200".replace(" ", ""))


def test_doc_example_encodes_byte_exact():
    assert G.encode_struct("Point", [("X", "int", 22), ("Y", "int", 33)]) == DOC_EXAMPLE


def test_doc_example_decodes():
    assert G.decode_stream(DOC_EXAMPLE) == [("Point", {"X": 22, "Y": 33})]


def test_uint_and_int_forms():
    assert G.enc_uint(0) == b"\x00" and G.enc_uint(127) == b"\x7f"
    assert G.enc_uint(256) == b"\xfe\x01\x00"  # documented example
    assert G.enc_int(-65) == b"\xff\x81" and G.enc_int(65) == b"\xff\x82"
    for v in (0, 1, -1, 63, 64, -64, -65, 1 << 40, -(1 << 40), (1 << 62), -(1 << 62)):
        assert G.Reader(G.enc_int(v)).int() == v
    for v in (0, 127, 128, 255, 256, 1 << 32, (1 << 64) - 1):
        assert G.Reader(G.enc_uint(v)).uint() == v


def test_ciphertext_roundtrip_and_zero_fields():
    c = bytes(range(1, 35))
    for l2 in (False, True):
        env = G.encode_ciphertext(c, l2)
        assert G.decode_ciphertext(env) == (c, l2)
    # the false flag is a zero value and is not transmitted: the value message is id, field 0, end
    env = G.encode_ciphertext(c, False)
    assert env.endswith(G.enc_uint(2 + 1 + 1 + len(c) + 1) + b"\xff\x82" + b"\x01" + G.enc_uint(len(c)) + c + b"\x00")
    assert G.decode_ciphertext(G.encode_ciphertext(b"", False)) == (b"", False)


def test_poly_ciphertext_roundtrip():
    coeffs = [bytes([i]) * 130 for i in range(1, 12)]
    for deg, sf, l2 in ((11, 0, False), (11, 3, True), (0, 0, False)):
        cs = coeffs[:deg] if deg else []
        assert G.decode_poly_ciphertext(G.encode_poly_ciphertext(cs, deg, sf, l2)) == (cs, deg, sf, l2)
    env = G.encode_poly_ciphertext(coeffs, 11, 2, True)
    names = [n for n, _ in G.decode_stream(env)]
    assert names == ["polyCiphertextWrapper"]


def test_decoder_accepts_foreign_ids_and_field_order():
    """Go assigns type ids per process and matches fields by name: a stream whose struct is id 70,
    whose slice is id 71 and whose fields come in another order must decode the same."""
    types = {70: G.StructT("polyCiphertextWrapper", [("L2", G.T_BOOL), ("ScaleFactor", G.T_INT),
                                                      ("CoeffBytes", 71), ("Degree", G.T_INT)]),
             71: G.SliceT("[][]uint8", G.T_BYTES)}
    val = {"L2": True, "ScaleFactor": 4, "CoeffBytes": [b"ab", b"cd"], "Degree": 2}
    stream = b"".join(G._message(G.enc_int(-t) + G._enc_wiretype(t, types[t])) for t in (70, 71))
    stream += G._message(G.enc_int(70) + G._enc_value(70, val, types))
    assert G.decode_poly_ciphertext(stream) == ([b"ab", b"cd"], 2, 4, True)


def test_malformed_streams_raise():
    env = G.encode_ciphertext(b"abc", True)
    for bad in (env[:-3], env[:5], b"\x03\xff\x82\x00"):
        with pytest.raises((G.GobError, KeyError)):
            G.decode_ciphertext(bad)


def test_float_and_bigint_forms():
    # documented: 17.0 is fe 31 40 (exponent and high mantissa only)
    assert G.enc_float(17.0) == bytes.fromhex("fe3140")
    for f in (0.0001, 1.5, -2.25, 1e300):
        assert G.dec_float(G.Reader(G.enc_float(f))) == f
    assert G.bigint_gob(0) == b"\x02" and G.bigint_gob(255) == b"\x02\xff" and G.bigint_gob(-256) == b"\x03\x01\x00"
    for v in (0, 1, -1, 1 << 512, -(1 << 100) + 7):
        assert G.bigint_ungob(G.bigint_gob(v)) == v


def test_public_key_roundtrip():
    n = (1 << 511) + 12345
    env = G.encode_public_key(b"g" * 130, b"p" * 130, b"q" * 130, n, 1 << 20, "type a1\np 7\nn 3\nl 4\n", True, 3, 3, 0.0001)
    w = G.decode_public_key(env)
    assert w == {"G1": b"g" * 130, "P": b"p" * 130, "Q": b"q" * 130, "N": n, "MsgSpace": 1 << 20,
                 "PairingParams": "type a1\np 7\nn 3\nl 4\n", "Deterministic": True, "PolyBase": 3, "FPScaleBase": 3,
                 "FPPrecision": 0.0001}
    names = [nm for nm, _ in G.decode_stream(env)]
    assert names == ["publicKeyWrapper"]
    # Deterministic=false is a zero value: omitted on the wire, false after decoding
    assert G.decode_public_key(G.encode_public_key(b"a", b"b", b"c", 5, 7, "x", False, 3, 3, 0.5))["Deterministic"] is False


def test_fuzz_roundtrips():
    """random envelopes (lengths around the 1-byte/multi-byte uint boundary, empty slices, large ints)"""
    from hypothesis import given, settings, strategies as st

    blob = st.binary(min_size=0, max_size=300)

    @settings(max_examples=150, deadline=None)
    @given(st.lists(blob, max_size=12), st.integers(0, 1 << 40), st.integers(-(1 << 40), 1 << 40), st.booleans())
    def poly(coeffs, deg, sf, l2):
        assert G.decode_poly_ciphertext(G.encode_poly_ciphertext(coeffs, deg, sf, l2)) == (coeffs, deg, sf, l2)

    @settings(max_examples=150, deadline=None)
    @given(blob, st.booleans())
    def single(c, l2):
        assert G.decode_ciphertext(G.encode_ciphertext(c, l2)) == (c, l2)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 1 << 1030), st.integers(0, 1 << 64), st.text(max_size=40), st.booleans(),
           st.integers(0, 100), st.floats(allow_nan=False, allow_infinity=False))
    def key(n, t, params, det, base, prec):
        w = G.decode_public_key(G.encode_public_key(b"g", b"p", b"q", n, t, params, det, base, base + 1, prec))
        assert (w["N"], w["MsgSpace"], w["PairingParams"], w["Deterministic"], w["PolyBase"], w["FPScaleBase"],
                w["FPPrecision"]) == (n, t, params, det, base, base + 1, prec)

    poly()
    single()
    key()
