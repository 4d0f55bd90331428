#!/usr/bin/env python
"""BASELINE.json config 5: fixed-point encrypted inner product sum_i u_i * v_i, sharded by index over
the GPUs of one box (one process per GPU, torchrun), checked against the plaintext result.

  per rank : Encrypt its shard of u and v (d coefficient slots each), MultPoly (d*d pairings per
             term), per-GPU GT product tree (bgn_l2_sum_reduce)                     -- no exchange
  exchange : ONE all-gather of 2d serialised GT elements per rank (NCCL; a few KB), then the same
             device reduction folds the world's partials on every rank
  check    : rank 0 decrypts the 2d result slots (T = 2^20) and compares them with the plaintext
             convolution sum computed independently with torch integer arithmetic

usage (N GPUs):  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
                     tools/inner_product.py [--key-bits 1024] [--length 65536] [--slots 8]
Prints one JSON object on rank 0."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bgn_b200 import Engine  # noqa: E402
from bgn_b200.multi import fold_l2_sum, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--key-bits", type=int, default=1024)
    ap.add_argument("--length", type=int, default=1 << 16)
    ap.add_argument("--slots", dest="d", type=int, default=8)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % args.key_bits)) as f:
        g = json.load(f)
    p, n, l, q1 = int(g["p"], 16), int(g["n"], 16), g["l"], int(g["q1"], 16)
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=local)
    d, SB = args.d, eng.scalar_bytes
    lo, hi = shard_range(args.length, rank, world)
    cnt = hi - lo
    gen = torch.Generator(device=dev)
    gen.manual_seed(5000 + rank)
    u = torch.randint(-1, 2, (cnt, d), generator=gen, device=dev, dtype=torch.int64)
    v = torch.randint(-1, 2, (cnt, d), generator=gen, device=dev, dtype=torch.int64)

    def rnd():
        r = torch.randint(0, 256, (cnt * d, SB), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        return r.reshape(-1)

    eng.timing_enable(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cu = eng.encrypt_batch(u.reshape(-1), rnd())
    cv = eng.encrypt_batch(v.reshape(-1), rnd())
    t_enc = eng.timing_last_call() * 2
    prod = eng.multpoly_batch(cu, d, cv, d, cnt)
    t_mult = eng.timing_last_call()
    part = eng.l2_sum_reduce(prod, cnt, 2 * d)
    t_red = eng.timing_last_call()
    total = fold_l2_sum(part, 2 * d, eng.l2_sum_reduce)  # device to device: NCCL all-gather + fold kernel
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0

    # plaintext reference: sum of per-term convolutions, with torch integer arithmetic
    conv = torch.zeros(2 * d, dtype=torch.int64, device=dev)
    for i in range(d):
        for k in range(d):
            conv[i + k] += (u[:, i] * v[:, k]).sum()
    times = torch.tensor([t_enc, t_mult, t_red, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(conv, op=dist.ReduceOp.SUM)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    if rank == 0:
        eng.set_secret(q1, 1 << 20)
        vals, status = eng.decrypt_batch(total, True)
        ok = (not status.any()) and [int(x) for x in vals] == [int(x) for x in conv.cpu()]
        t_enc, t_mult, t_red, wall_ms = [float(x) for x in times.cpu()]
        print(json.dumps({
            "config": "keyBits=%d encrypted inner product, length %d, d=%d slots, %d GPU(s)" % (
                args.key_bits, args.length, d, world),
            "decrypted_matches_plaintext": bool(ok), "slots": [int(x) for x in vals],
            "emult_per_s": args.length / (t_mult * 1e-3), "pairings_per_s": args.length * d * d / (t_mult * 1e-3),
            "ms_max_over_ranks": {"encrypt": t_enc, "multpoly": t_mult, "l2_sum_tree": t_red, "wall": wall_ms},
            "exchange_bytes_per_rank": int(part.numel()), "n_gpus": world}))
        assert ok, "inner product mismatch"
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
