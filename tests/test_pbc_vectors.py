"""PBC-made known answers: when tests/golden/pbc/kb*.json exists -- written by
tools/pbc_vectors/main.go with the real libpbc -- every output it holds must equal the committed,
oracle-derived fixture byte for byte.  That turns "parity by definition" into "parity by KAT"."""
import glob
import json
import os

import pytest

from conftest import GOLDEN_DIR, load_golden

FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "pbc", "kb*.json")))


@pytest.mark.skipif(not FILES, reason="no PBC-produced vectors (needs Go + libpbc: tools/pbc_vectors/main.go)")
@pytest.mark.parametrize("path", FILES or ["-"], ids=os.path.basename)
def test_pbc_vectors_match_fixture(path):
    with open(path) as f:
        pbc = json.load(f)
    g = load_golden(int(pbc["key_bits"]))
    checked = 0
    gobs = pbc.pop("gob", None)
    if gobs:  # envelopes written by Go's encoding/gob: the Python codec must read them, and write the same bytes
        from bgn_b200 import gobwire
        c1 = [bytes.fromhex(h) for h in g["multpoly"]["c1"]]
        pair0 = bytes.fromhex(g["pair"]["out"][0])
        assert gobwire.decode_ciphertext(bytes.fromhex(gobs["ciphertext_l1"])) == (c1[0], False)
        assert gobwire.decode_ciphertext(bytes.fromhex(gobs["ciphertext_l2"])) == (pair0, True)
        assert gobwire.decode_poly_ciphertext(bytes.fromhex(gobs["poly_ciphertext"])) == (c1, len(c1), 2, False)
        # a fresh Go process numbers its first user type 65, as the Python encoder does
        assert gobwire.encode_ciphertext(c1[0], False) == bytes.fromhex(gobs["ciphertext_l1"])
        checked += 3
    for name, sec in pbc.items():
        if not isinstance(sec, dict):
            continue
        for key, val in sec.items():
            assert g[name][key] == val, "%s.%s differs from libpbc" % (name, key)
            checked += len(val)
    assert checked > 0


def test_schema_of_the_dump_tool_matches_fixture():
    """every section/key the Go tool writes exists in the fixtures (keeps the tool and the fixtures in step)"""
    src = open(os.path.join(os.path.dirname(GOLDEN_DIR), "..", "tools", "pbc_vectors", "main.go")).read()
    g = load_golden(64)
    for name in ("encrypt", "g1_add", "g1_sub", "g1_neg", "g1_mulconst", "pair", "make_l2", "gt_mul", "gt_div", "gt_inv",
                 "gt_pow", "multpoly", "decrypt_l2", "g1_blind", "gt_blind"):
        assert '"%s"' % name in src and name in g
    assert "csk" in g["decrypt_l2"]
