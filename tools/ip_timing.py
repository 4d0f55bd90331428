#!/usr/bin/env python
"""Where the time of a keyBits=1024 MultPoly call (config 5 shape: d = 8) goes: device ms of the call and of each kernel
class, first and second call."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bgn_b200 import Engine

def main():
    cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb1024.json")) as f:
        g = json.load(f)
    eng = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    d = 8
    def enc(n):
        x = torch.randint(-1, 2, (n,), generator=gen, device=dev, dtype=torch.int64)
        r = torch.randint(0, 256, (n, eng.scalar_bytes), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        return eng.encrypt_batch(x, r.reshape(-1))
    a, b = enc(cnt * d), enc(cnt * d)
    eng.timing_enable(True)
    for it in range(3):
        eng.timing_reset()
        out = eng.multpoly_batch(a, d, b, d, cnt)
        row = {"call": it, "call_ms": eng.timing_last_call()}
        for k in ("k_miller", "k_g1_from_bytes", "k_fp2_to_bytes", "k_normalize", "k_"):
            row[k] = eng.timing_get(k)
        print(json.dumps(row))
    eng.close()

main()
