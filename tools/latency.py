#!/usr/bin/env python
"""BASELINE.json config 1 (`go test -bench=.` analogue, bgn_test.go:97-140): single-item Encrypt(1),
Add(c, c), Mult(c, c), MultConst(c, 1) and Decrypt through the host mirror at keyBits 128 and 512 --
a batch of ONE per call, host buffers, wall-clock per call (median of `reps`).  The engine is built
for batches: a single pairing is one GPU thread walking ~13 000 dependent Montgomery products, so
these latencies are the price of the design, reported for completeness (the CPU port of the oracle
needs ~2 ms per 512-bit pairing on one host core: bench.py's cpu_baseline).  Prints one JSON object.   usage: tools/latency.py [--reps 20]"""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgn_b200 import PublicKey, SecretKey  # noqa: E402


def med_ms(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    out = {}
    for kb in (128, 512):
        with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % kb)) as f:
            g = json.load(f)
        pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), 1021)
        sk = SecretKey(int(g["q1"], 16))
        pk.SetupDecryption(sk)
        c = pk.Encrypt(1)
        m = pk.Mult(c, c)
        res = {
            "Encrypt": med_ms(lambda: pk.Encrypt(1), args.reps),
            "Add": med_ms(lambda: pk.Add(c, c), args.reps),
            "Mult": med_ms(lambda: pk.Mult(c, c), args.reps),
            "MultConst": med_ms(lambda: pk.MultConst(c, 1), args.reps),
            "Decrypt_L1": med_ms(lambda: sk.Decrypt(c, pk), args.reps),
            "Decrypt_L2": med_ms(lambda: sk.Decrypt(m, pk), args.reps),
        }
        assert sk.Decrypt(pk.Add(c, c), pk) == 2 and sk.Decrypt(m, pk) == 1
        out["kb%d" % kb] = {"gpu_ms_per_call_batch_of_1": res}
        pk.engine.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
