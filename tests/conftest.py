import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
KEY_BITS = (64, 128, 256, 512, 1024)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(kb: int) -> dict:
    with open(os.path.join(GOLDEN_DIR, "kb%d.json" % kb)) as f:
        return json.load(f)


@pytest.fixture(scope="session", params=KEY_BITS, ids=lambda kb: "kb%d" % kb)
def golden(request):
    return load_golden(request.param)


def unhex(xs):
    return b"".join(bytes.fromhex(x) for x in xs)
