"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard boundaries and the
all-gather + fold of per-rank L2 partial sums.  The device reduction is replaced by the oracle's
GT product here (tests may use the oracle); the GPU suite checks Engine.l2_sum_reduce itself."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden, unhex


def test_shard_range():
    from bgn_b200.multi import shard_range
    for count in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(count, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == count
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from bgn_b200.multi import fold_l2_sum, gather_partials
    from oracle import bgn_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(64)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    v = g["l2_sum"]
    ncoeff, EB = v["ncoeff"], 2 * par.coord_bytes
    terms = np.frombuffer(unhex(v["in"]), dtype=np.uint8).reshape(v["nterms"], ncoeff * EB)
    mine = terms[rank::world]  # this rank's share of the terms

    def cpu_reduce(buf, nterms, nc):  # stands in for Engine.l2_sum_reduce on the CPU
        out = []
        for c in range(nc):
            acc = O.GT_ONE
            for t in range(nterms):
                e = bytes(buf[(t * nc + c) * EB:(t * nc + c + 1) * EB])
                acc = O.fp2_mul(acc, O.gt_from_bytes(e, par), par.p)
            out.append(O.gt_to_bytes(acc, par))
        return np.frombuffer(b"".join(out), dtype=np.uint8)

    part = cpu_reduce(mine.reshape(-1), mine.shape[0], ncoeff)
    allp = gather_partials(part)
    total = fold_l2_sum(part, ncoeff, cpu_reduce)
    q.put((rank, allp.shape, bytes(allp[rank]) == part.tobytes(), total.tobytes() == unhex(v["out"])))
    dist.destroy_process_group()


def test_fold_l2_sum_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, shape, own_ok, total_ok in res:
        assert shape[0] == 2 and own_ok and total_ok


def test_fold_single_process_passthrough():
    from bgn_b200.multi import fold_l2_sum
    buf = np.arange(40, dtype=np.uint8)
    assert fold_l2_sum(buf, 2, lambda *a: pytest.fail("no reduction needed")).tobytes() == buf.tobytes()
