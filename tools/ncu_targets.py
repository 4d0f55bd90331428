#!/usr/bin/env python
"""One call of one operation on cuda:0, for `ncu -k regex:<kernel> -c 1` captures (tools/gpu_r2h.sh).
usage: tools/ncu_targets.py <fixed_pair|pair_duo|split|miller|miller1024|dec_lucas|encrypt> [count]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bgn_b200 import Engine  # noqa: E402


def main():
    which = sys.argv[1]
    kb = 1024 if which == "miller1024" else 512
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % kb)) as f:
        g = json.load(f)
    eng = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    EB, SB = eng.elem_bytes, eng.scalar_bytes
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    default = {"fixed_pair": 1 << 14, "pair_duo": 1 << 14, "split": 2048, "miller": 3404, "miller1024": 2368, "dec_lucas": 1 << 14, "encrypt": 720896}
    cnt = int(sys.argv[2]) if len(sys.argv) > 2 else default[which]
    d = {"split": 11, "miller": 11, "miller1024": 8}.get(which, 1)

    def enc(n):
        x = torch.randint(-1, 2, (n,), generator=gen, device=dev, dtype=torch.int64)
        r = torch.randint(0, 256, (n, SB), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        return eng.encrypt_batch(x, r.reshape(-1))

    if which == "encrypt":  # warm call builds the table, the second launch of k_encrypt is the one to capture
        enc(4096)
        enc(cnt)
        torch.cuda.synchronize()
        eng.close()
        print("done", which, cnt)
        return
    a, b = enc(cnt * d), enc(cnt * d)
    if which == "fixed_pair":
        eng.make_l2_batch(a)
    elif which == "pair_duo":
        eng.pair_batch(a, b)
    elif which in ("split", "miller", "miller1024"):
        eng.set_option("miller_split", 1 if which == "split" else 0)
        eng.multpoly_batch(a, d, b, d, cnt)
    elif which == "dec_lucas":
        eng.set_secret(int(g["q1"], 16), 1 << 20)
        l2 = eng.pair_batch(a, b)
        eng.decrypt_batch(l2, True)
    torch.cuda.synchronize()
    eng.close()
    print("done", which, cnt)


if __name__ == "__main__":
    main()
