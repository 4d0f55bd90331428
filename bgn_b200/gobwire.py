"""encoding/gob envelopes of the reference's ciphertext types.

The reference serialises ciphertexts as gob streams of two private wrapper structs
(ciphertext.go:17-20, 33-38, 76-116) and parses them back in NewCiphertextFromBytes /
NewPolyCiphertextFromBytes (bgn.go:505-555):

    type ciphertextWrapper struct     { CBytes []byte;  L2 bool }
    type polyCiphertextWrapper struct { CoeffBytes [][]byte; Degree int; ScaleFactor int; L2 bool }

This module restates the gob wire format (the Go documentation of package encoding/gob is the
specification) for exactly the kinds those two structs need -- structs, slices, []byte, string, bool,
int, uint -- so that the host mirror can read what a Go process wrote and write what a Go process
reads, with no Go toolchain.  The decoder is general within those kinds: it accepts any type ids
(ids are assigned per process in registration order, so they are not a constant of the format) and
matches struct fields BY NAME, as gob does.  The encoder numbers its types from 65, the first user id
of a fresh process.

Verification status: checked against the worked example of the gob documentation (the `Point{22, 33}`
stream, tests/test_gobwire.py); there is no Go here to produce envelopes of the bgn wrappers
themselves -- tools/pbc_vectors can be extended to dump them where Go exists.

Unlike the reference's decoder, `decode_*` never evaluates a pairing: the reference computes e(Q,Q)
once per level-2 coefficient merely to obtain a GT-typed element to SetBytes into (bgn.go:517, 549);
here the element bytes go to the device as they are.
"""
from __future__ import annotations

import struct
from typing import Any, Dict, List, Optional, Tuple

# builtin type ids (encoding/gob type.go)
T_BOOL, T_INT, T_UINT, T_FLOAT, T_BYTES, T_STRING, T_COMPLEX, T_INTERFACE = 1, 2, 3, 4, 5, 6, 7, 8
T_WIRETYPE, T_ARRAYTYPE, T_COMMONTYPE, T_SLICETYPE, T_STRUCTTYPE, T_FIELDTYPE, T_FIELDSLICE, T_MAPTYPE = (
    16, 17, 18, 19, 20, 21, 22, 23)
FIRST_USER_ID = 65


class GobError(ValueError):
    pass


# ----------------------------------------------------------------------------- primitives
def enc_uint(u: int) -> bytes:
    """< 128: one byte; else the negated byte count followed by the big-endian value."""
    if u < 0:
        raise GobError("negative uint")
    if u < 128:
        return bytes([u])
    raw = u.to_bytes((u.bit_length() + 7) // 8, "big")
    return bytes([256 - len(raw)]) + raw


def enc_int(i: int) -> bytes:
    """bit 0 = complement flag, value in the upper bits."""
    return enc_uint((~i << 1) | 1 if i < 0 else i << 1)


def enc_bytes(b: bytes) -> bytes:
    return enc_uint(len(b)) + bytes(b)


class Reader:
    def __init__(self, data: bytes, pos: int = 0, end: Optional[int] = None):
        self.d, self.pos, self.end = data, pos, len(data) if end is None else end

    def eof(self) -> bool:
        return self.pos >= self.end

    def take(self, n: int) -> bytes:
        if n < 0 or self.pos + n > self.end:
            raise GobError("truncated gob stream")
        out = self.d[self.pos:self.pos + n]
        self.pos += n
        return out

    def uint(self) -> int:
        b = self.take(1)[0]
        if b < 128:
            return b
        n = 256 - b
        if n > 8:
            raise GobError("uint wider than 64 bits")
        return int.from_bytes(self.take(n), "big")

    def int(self) -> int:
        u = self.uint()
        return ~(u >> 1) if u & 1 else u >> 1

    def bytes_(self) -> bytes:
        return self.take(self.uint())


# ----------------------------------------------------------------------------- type descriptions
class SliceT:
    def __init__(self, name: str, elem: int):
        self.name, self.elem = name, elem


class StructT:
    def __init__(self, name: str, fields: List[Tuple[str, int]]):
        self.name, self.fields = name, fields


class GobEncT:
    """a type that implements GobEncoder (math/big.Int here): its values travel as opaque bytes"""

    def __init__(self, name: str):
        self.name = name


T_GOBENCTYPE = 1000  # gobEncoderType{CommonType}: no fixed id in Go, only its shape matters here


def enc_float(f: float) -> bytes:
    """float64 bits, byte-reversed (small exponents and mantissas with trailing zeros become short uints)"""
    return enc_uint(int.from_bytes(struct.pack("<d", float(f)), "big"))


def dec_float(r: "Reader") -> float:
    return struct.unpack("<d", r.uint().to_bytes(8, "big"))[0]


def bigint_gob(x: int) -> bytes:
    """math/big.(*Int).GobEncode: version 1 in the upper 7 bits, sign in bit 0, then the magnitude"""
    mag = abs(x)
    return bytes([(1 << 1) | (1 if x < 0 else 0)]) + mag.to_bytes((mag.bit_length() + 7) // 8, "big")


def bigint_ungob(b: bytes) -> int:
    if len(b) == 0:
        return 0
    if b[0] >> 1 != 1:
        raise GobError("unsupported big.Int gob version %d" % (b[0] >> 1))
    v = int.from_bytes(b[1:], "big")
    return -v if b[0] & 1 else v


def _go_name(tid: int, types: Dict[int, Any]) -> str:
    return {T_BOOL: "bool", T_INT: "int", T_UINT: "uint", T_FLOAT: "float64", T_BYTES: "[]uint8",
            T_STRING: "string"}.get(tid) or types[tid].name


def _enc_common(name: str, tid: int) -> bytes:
    # CommonType{Name string; Id typeId}
    return b"\x01" + enc_bytes(name.encode()) + b"\x01" + enc_int(tid) + b"\x00"


def _enc_wiretype(tid: int, t: Any) -> bytes:
    """a value of wireType{ArrayT, SliceT, StructT, MapT, ...} describing user type `tid`"""
    if isinstance(t, SliceT):
        body = b"\x01" + _enc_common(t.name, tid) + b"\x01" + enc_int(t.elem) + b"\x00"  # sliceType
        return b"\x02" + body + b"\x00"  # field 1 of wireType
    if isinstance(t, GobEncT):
        return b"\x05" + b"\x01" + _enc_common(t.name, tid) + b"\x00" + b"\x00"  # field 4 of wireType
    body = b"\x01" + _enc_common(t.name, tid)
    if t.fields:
        body += b"\x01" + enc_uint(len(t.fields))
        for fname, ftid in t.fields:
            body += b"\x01" + enc_bytes(fname.encode()) + b"\x01" + enc_int(ftid) + b"\x00"  # fieldType
    body += b"\x00"
    return b"\x03" + body + b"\x00"  # field 2 of wireType


def _message(payload: bytes) -> bytes:
    return enc_uint(len(payload)) + payload


# ----------------------------------------------------------------------------- values
def _is_zero(tid: int, v: Any, types) -> bool:
    if tid in (T_BOOL,):
        return not v
    if tid in (T_INT, T_UINT):
        return v == 0
    if tid in (T_BYTES, T_STRING):
        return len(v) == 0
    if tid == T_FLOAT:
        return v == 0.0
    t = types[tid]
    if isinstance(t, SliceT):
        return len(v) == 0
    return v is None  # nil pointers (structs, GobEncoders) are not transmitted


def _enc_value(tid: int, v: Any, types) -> bytes:
    if tid == T_BOOL:
        return enc_uint(1 if v else 0)
    if tid == T_INT:
        return enc_int(int(v))
    if tid == T_UINT:
        return enc_uint(int(v))
    if tid == T_BYTES:
        return enc_bytes(bytes(v))
    if tid == T_STRING:
        return enc_bytes(v.encode() if isinstance(v, str) else bytes(v))
    if tid == T_FLOAT:
        return enc_float(v)
    t = types[tid]
    if isinstance(t, SliceT):
        return enc_uint(len(v)) + b"".join(_enc_value(t.elem, x, types) for x in v)
    if isinstance(t, GobEncT):
        return enc_bytes(v)
    out, last = b"", -1
    for idx, (fname, ftid) in enumerate(t.fields):
        fv = v[fname]
        if _is_zero(ftid, fv, types):  # zero values are not transmitted
            continue
        out += enc_uint(idx - last) + _enc_value(ftid, fv, types)
        last = idx
    return out + b"\x00"


def _dec_value(r: Reader, tid: int, types) -> Any:
    if tid == T_BOOL:
        return r.uint() != 0
    if tid == T_INT:
        return r.int()
    if tid == T_UINT:
        return r.uint()
    if tid == T_BYTES:
        return r.bytes_()
    if tid == T_STRING:
        return r.bytes_().decode()
    if tid == T_FLOAT:
        return dec_float(r)
    if tid not in types:
        raise GobError("value of undefined type id %d" % tid)
    t = types[tid]
    if isinstance(t, GobEncT):
        return r.bytes_()
    if isinstance(t, SliceT):
        n = r.uint()
        if n > r.end - r.pos:
            raise GobError("slice longer than the message")
        return [_dec_value(r, t.elem, types) for _ in range(n)]
    out: Dict[str, Any] = {}
    idx = -1
    while True:
        delta = r.uint()
        if delta == 0:
            return out
        idx += delta
        if idx >= len(t.fields):
            raise GobError("field number out of range for %s" % t.name)
        fname, ftid = t.fields[idx]
        out[fname] = _dec_value(r, ftid, types)


# the bootstrap types needed to read a wireType value
_BOOT: Dict[int, Any] = {
    T_COMMONTYPE: StructT("CommonType", [("Name", T_STRING), ("Id", T_INT)]),
    T_ARRAYTYPE: StructT("arrayType", [("CommonType", T_COMMONTYPE), ("Elem", T_INT), ("Len", T_INT)]),
    T_SLICETYPE: StructT("sliceType", [("CommonType", T_COMMONTYPE), ("Elem", T_INT)]),
    T_FIELDTYPE: StructT("fieldType", [("Name", T_STRING), ("Id", T_INT)]),
    T_FIELDSLICE: SliceT("[]*gob.fieldType", T_FIELDTYPE),
    T_STRUCTTYPE: StructT("structType", [("CommonType", T_COMMONTYPE), ("Field", T_FIELDSLICE)]),
    T_MAPTYPE: StructT("mapType", [("CommonType", T_COMMONTYPE), ("Key", T_INT), ("Elem", T_INT)]),
    T_GOBENCTYPE: StructT("gobEncoderType", [("CommonType", T_COMMONTYPE)]),
    T_WIRETYPE: StructT("wireType", [("ArrayT", T_ARRAYTYPE), ("SliceT", T_SLICETYPE), ("StructT", T_STRUCTTYPE),
                                     ("MapT", T_MAPTYPE), ("GobEncoderT", T_GOBENCTYPE),
                                     ("BinaryMarshalerT", T_GOBENCTYPE), ("TextMarshalerT", T_GOBENCTYPE)]),
}


def decode_stream(data: bytes) -> List[Tuple[str, Any]]:
    """-> [(type name, value)] for every value message of the stream (type definitions are consumed)."""
    r = Reader(bytes(data))
    types: Dict[int, Any] = {}
    out = []
    while not r.eof():
        n = r.uint()
        m = Reader(r.d, r.pos, r.pos + n)
        r.take(n)
        tid = m.int()
        if tid < 0:  # type definition
            w = _dec_value(m, T_WIRETYPE, _BOOT)
            if "StructT" in w:
                st = w["StructT"]
                types[-tid] = StructT(st.get("CommonType", {}).get("Name", ""),
                                      [(f.get("Name", ""), f.get("Id", 0)) for f in st.get("Field", [])])
            elif "SliceT" in w:
                sl = w["SliceT"]
                types[-tid] = SliceT(sl.get("CommonType", {}).get("Name", ""), sl.get("Elem", 0))
            elif "GobEncoderT" in w or "BinaryMarshalerT" in w:
                ge = w.get("GobEncoderT") or w.get("BinaryMarshalerT")
                types[-tid] = GobEncT(ge.get("CommonType", {}).get("Name", ""))
            else:
                raise GobError("unsupported wire type (only structs and slices occur in the bgn envelopes)")
            continue
        if tid in types and isinstance(types[tid], StructT):
            val = _dec_value(m, tid, types)
        else:  # non-struct top-level values carry a zero "singleton" delta byte
            if m.uint() != 0:
                raise GobError("corrupt singleton value")
            val = _dec_value(m, tid, types)
        if not m.eof():
            raise GobError("trailing bytes in a gob message")
        out.append((_go_name(tid, types), val))
    return out


def encode_struct(name: str, fields: List[Tuple[str, Any, Any]]) -> bytes:
    """One gob stream holding one struct value.  fields: (name, type, value) with type one of
    'bool', 'int', 'uint', 'bytes', 'string', '[][]byte'.  Type definitions are sent the way
    Encoder.sendType does: the struct first, then the types of its fields."""
    types: Dict[int, Any] = {}
    sid = FIRST_USER_ID
    nxt = sid + 1
    flds: List[Tuple[str, int]] = []
    value: Dict[str, Any] = {}
    for fname, ftype, fval in fields:
        if ftype == "[][]byte":
            ftid = next((k for k, t in types.items() if isinstance(t, SliceT) and t.elem == T_BYTES), None)
            if ftid is None:
                ftid = nxt
                nxt += 1
                types[ftid] = SliceT("[][]uint8", T_BYTES)
        else:
            ftid = {"bool": T_BOOL, "int": T_INT, "uint": T_UINT, "bytes": T_BYTES, "string": T_STRING}[ftype]
        flds.append((fname, ftid))
        value[fname] = fval
    types[sid] = StructT(name, flds)
    out = _message(enc_int(-sid) + _enc_wiretype(sid, types[sid]))
    for tid in sorted(k for k in types if k != sid):
        out += _message(enc_int(-tid) + _enc_wiretype(tid, types[tid]))
    return out + _message(enc_int(sid) + _enc_value(sid, value, types))


# ----------------------------------------------------------------------------- the bgn envelopes
def encode_ciphertext(C: bytes, L2: bool) -> bytes:
    """Ciphertext.Bytes() (ciphertext.go:76-91)."""
    return encode_struct("ciphertextWrapper", [("CBytes", "bytes", C), ("L2", "bool", L2)])


def decode_ciphertext(data: bytes) -> Tuple[bytes, bool]:
    """NewCiphertextFromBytes (bgn.go:505-528) without the per-call pairing."""
    vals = decode_stream(data)
    if len(vals) != 1:
        raise GobError("expected one value")
    v = vals[0][1]
    return bytes(v.get("CBytes", b"")), bool(v.get("L2", False))


def encode_poly_ciphertext(coeffs: List[bytes], Degree: int, ScaleFactor: int, L2: bool) -> bytes:
    """PolyCiphertext.Bytes() (ciphertext.go:93-116)."""
    return encode_struct("polyCiphertextWrapper", [("CoeffBytes", "[][]byte", list(coeffs)), ("Degree", "int", Degree),
                                                   ("ScaleFactor", "int", ScaleFactor), ("L2", "bool", L2)])


def decode_poly_ciphertext(data: bytes) -> Tuple[List[bytes], int, int, bool]:
    """NewPolyCiphertextFromBytes (bgn.go:530-555)."""
    vals = decode_stream(data)
    if len(vals) != 1:
        raise GobError("expected one value")
    v = vals[0][1]
    return ([bytes(c) for c in v.get("CoeffBytes", [])], int(v.get("Degree", 0)), int(v.get("ScaleFactor", 0)),
            bool(v.get("L2", False)))


def encode_public_key(G1: bytes, P: bytes, Q: bytes, N: int, MsgSpace: int, PairingParams: str, Deterministic: bool,
                      PolyBase: int, FPScaleBase: int, FPPrecision: float) -> bytes:
    """PublicKey.MarshalBinary (bgn.go:597-624): gob of publicKeyWrapper{G1, P, Q []byte; N, MsgSpace *big.Int;
    PairingParams string; Deterministic bool; PolyEncodingParams *PolyEncodingParams} (bgn.go:45-55)."""
    sid, bid, pid = FIRST_USER_ID, FIRST_USER_ID + 1, FIRST_USER_ID + 2
    types = {
        sid: StructT("publicKeyWrapper", [("G1", T_BYTES), ("P", T_BYTES), ("Q", T_BYTES), ("N", bid), ("MsgSpace", bid),
                                          ("PairingParams", T_STRING), ("Deterministic", T_BOOL),
                                          ("PolyEncodingParams", pid)]),
        bid: GobEncT("Int"),
        pid: StructT("PolyEncodingParams", [("PolyBase", T_INT), ("FPScaleBase", T_INT), ("FPPrecision", T_FLOAT)]),
    }
    value = {"G1": G1, "P": P, "Q": Q, "N": bigint_gob(N), "MsgSpace": bigint_gob(MsgSpace),
             "PairingParams": PairingParams, "Deterministic": Deterministic,
             "PolyEncodingParams": {"PolyBase": PolyBase, "FPScaleBase": FPScaleBase, "FPPrecision": FPPrecision}}
    out = b"".join(_message(enc_int(-t) + _enc_wiretype(t, types[t])) for t in (sid, bid, pid))
    return out + _message(enc_int(sid) + _enc_value(sid, value, types))


def decode_public_key(data: bytes) -> Dict[str, Any]:
    """PublicKey.UnmarshalBinary (bgn.go:628-666) up to, not including, the PBC objects: plain values."""
    vals = decode_stream(data)
    if len(vals) != 1:
        raise GobError("expected one value")
    v = vals[0][1]
    pe = v.get("PolyEncodingParams") or {}
    return {"G1": bytes(v.get("G1", b"")), "P": bytes(v.get("P", b"")), "Q": bytes(v.get("Q", b"")),
            "N": bigint_ungob(v.get("N", b"")), "MsgSpace": bigint_ungob(v.get("MsgSpace", b"")),
            "PairingParams": v.get("PairingParams", ""), "Deterministic": bool(v.get("Deterministic", False)),
            "PolyBase": int(pe.get("PolyBase", 0)), "FPScaleBase": int(pe.get("FPScaleBase", 0)),
            "FPPrecision": float(pe.get("FPPrecision", 0.0))}
