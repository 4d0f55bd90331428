"""The boundary from plain C (no Python, no torch): include/bgn_b200.h compiles as strict C11 and links;
on the GPU box a C program drives MultPoly -> L2 sum -> Decrypt through it and checks the golden bytes."""
import os
import subprocess

import pytest

from conftest import ROOT, load_golden

CABI = os.path.join(ROOT, "tests", "cabi")
LIBDIR = os.path.join(ROOT, "bgn_b200")
CC = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include")]
LINK = ["-L", LIBDIR, "-lbgn_b200", "-Wl,-rpath," + LIBDIR]


def test_header_is_plain_c_and_links(tmp_path):
    exe = str(tmp_path / "header_is_c")
    subprocess.check_call(CC + [os.path.join(CABI, "header_is_c.c"), "-o", exe] + LINK)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    n = int(out.stdout.split("/")[0])
    from bgn_b200 import _cabi
    assert n == len(_cabi.SIGNATURES)  # the C file lists every declared entry point


def c_array(name, data: bytes) -> str:
    return "static const uint8_t %s[] = {%s};\n" % (name, ",".join(str(b) for b in data))


@pytest.mark.gpu
def test_c_program_reproduces_golden_multpoly(tmp_path):
    g = load_golden(128)
    v = g["multpoly"]
    p, n, q1 = int(g["p"], 16), int(g["n"], 16), int(g["q1"], 16)
    be = lambda x: x.to_bytes((x.bit_length() + 7) // 8, "big")  # noqa: E731
    # plaintext digits of the golden polynomials (tests/golden/make_golden.py): [1, 0, 2] x [2, 1(O), 1, 0]
    d1, d2 = v["d1"], v["d2"]
    a, b = [1, 0, 2][:d1], [2, 0, 1, 0][:d2]  # c2[1] is O in the fixture: an encryption of 0
    plain = [0] * (d1 + d2)
    for i, x in enumerate(a):
        for k, y in enumerate(b):
            plain[i + k] += x * y
    hdr = "#include <stdint.h>\n#define D1 %d\n#define D2 %d\n#define KEY_L %dULL\n#define MSG_SPACE %dULL\n" % (
        d1, d2, g["l"], g["msg_space"])
    hdr += c_array("KEY_P", be(p)) + c_array("KEY_N", be(n)) + c_array("KEY_Q1", be(q1))
    hdr += c_array("KEY_GEN_P", bytes.fromhex(g["P"])) + c_array("KEY_GEN_Q", bytes.fromhex(g["Q"]))
    hdr += c_array("C1", b"".join(bytes.fromhex(h) for h in v["c1"]))
    hdr += c_array("C2", b"".join(bytes.fromhex(h) for h in v["c2"]))
    hdr += c_array("EXPECT", b"".join(bytes.fromhex(h) for h in v["out"]))
    hdr += "static const int64_t PLAIN[] = {%s};\n" % ",".join(str(x) for x in plain)
    (tmp_path / "vectors.h").write_text(hdr)
    exe = str(tmp_path / "emult_example")
    subprocess.check_call(CC + ["-I", str(tmp_path), os.path.join(CABI, "emult_example.c"), "-o", exe] + LINK)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "multpoly bytes == golden" in out.stdout


def write_key_header(path, g, d1, d2, msg_space=None):
    p, n, q1 = int(g["p"], 16), int(g["n"], 16), int(g["q1"], 16)
    be = lambda x: x.to_bytes((x.bit_length() + 7) // 8, "big")  # noqa: E731
    hdr = "#include <stdint.h>\n#define D1 %d\n#define D2 %d\n#define KEY_L %dULL\n#define MSG_SPACE %dULL\n" % (
        d1, d2, g["l"], msg_space or g["msg_space"])
    hdr += c_array("KEY_P", be(p)) + c_array("KEY_N", be(n)) + c_array("KEY_Q1", be(q1))
    hdr += c_array("KEY_GEN_P", bytes.fromhex(g["P"])) + c_array("KEY_GEN_Q", bytes.fromhex(g["Q"]))
    with open(path, "w") as f:
        f.write(hdr)


def build_multi_gpu(tmp_path, g, d1, d2, msg_space=None):
    write_key_header(str(tmp_path / "vectors.h"), g, d1, d2, msg_space)
    exe = str(tmp_path / "multi_gpu")
    subprocess.check_call(CC + ["-D_POSIX_C_SOURCE=200809L", "-I", str(tmp_path), os.path.join(CABI, "multi_gpu.c"), "-o", exe,
                                "-pthread"] + LINK)
    return exe


def test_multi_gpu_c_program_compiles(tmp_path):
    """the pthread driver of several contexts is strict C11 against the header (no GPU needed to build it)"""
    build_multi_gpu(tmp_path, load_golden(128), 3, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("kb,terms", [(128, 37), (512, 10)])
def test_c_threads_drive_several_contexts(tmp_path, kb, terms):
    """One process, one pthread per context: every GPU of the box when there are several (one context
    each), and two contexts sharing device 0 (serialised by the per-device lock) in any case.  The folded
    partial L2 sums decrypt to the plaintext inner product and equal the single-context bytes."""
    import torch
    exe = build_multi_gpu(tmp_path, load_golden(kb), 3, 4)
    ndev = torch.cuda.device_count()
    layouts = ["0,0", "0,0,0"]
    if ndev >= 2:
        layouts.append(",".join(str(i) for i in range(ndev)))
    for devs in layouts:
        out = subprocess.run([exe, devs, str(terms)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "bytes == single context" in out.stdout
