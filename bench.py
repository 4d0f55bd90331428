#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its headline config.

  metric   : reference-equivalent Tate pairings/s (= EMult/s x d1*d2) at 512-bit keys
  workload : config 3 -- batched EMult (MultPoly, poly.go:123-156) of 2^14 pairs of level-1
             polynomial ciphertexts with d1 = d2 = 11 coefficient slots (121 pairings and 22
             output slots per EMult); one "step" = one bgn_multpoly_batch call over the batch.
  N > 1    : every rank (one process per GPU) runs its own 2^14 pairs -- independent units, no
             data-path collective -> "scaling": "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

`value` is measured with inputs resident in HBM (CUDA events on the library's stream around each
call); `e2e` goes through the same C-ABI call with pinned HOST buffers, so the host<->device copies
are inside the timed region.  `roofline` is against the integer-multiply (IMAD.WIDE) pipe, which is
what bounds this path; HBM traffic is reported beside it as evidence that memory is
not the limiter.  `cpu_baseline` / `--impl reference` time the C port of the CPU oracle (oracle/cpu_ref.c; the
real reference needs Go + libpbc + GMP, none of which exist on the box) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D1 = D2 = 11
KEY_BITS = 512
# dram__bytes_read.sum + dram__bytes_write.sum of one k_miller<17> launch over 3404 units (one full
# wave), from the `ncu --set full` capture summarised in profiles/r01_miller_v24_ncu.txt
NCU_DRAM_BYTES_PER_UNIT = 10858752 / 3404.0
METRIC = "pairings/s"
WORKLOAD = "keyBits=512 batched EMult (MultPoly) of 2^14 L1 poly-ciphertext pairs, d1=d2=11 (121 pairings/EMult)"


def load_key():
    with open(os.path.join(ROOT, "tests", "golden", "kb%d.json" % KEY_BITS)) as f:
        return json.load(f)


# ----------------------------------------------------------------------------- CPU arm
def cpu_emults(count: int, cores: int):
    """`count` EMults (d1 = d2 = 11: 121 full Tate pairings + slot accumulation each) with the C port
    of the oracle (oracle/cpu_ref.c, 64-bit-limb Montgomery, one Miller loop + final exponentiation
    per pairing as libpbc does), `cores` threads.  -> (pairings/s, seconds)"""
    import numpy as np
    from oracle.cpu_ref import CpuRef
    g = load_key()
    ref = CpuRef(int(g["p"], 16), int(g["n"], 16), g["l"], threads=cores)
    rng = np.random.default_rng(7)
    x1 = rng.integers(-1, 2, count * D1)
    x2 = rng.integers(-1, 2, count * D2)
    # inputs: deterministic encryptions times a fixed blinding point, cheap to build on the CPU
    P, Q = bytes.fromhex(g["P"]), bytes.fromhex(g["Q"])
    r = np.zeros((count * D1, ref.nbytes), dtype=np.uint8)
    r[:, -2:] = rng.integers(1, 256, (count * D1, 2))
    c1 = ref.encrypt_batch(P, Q, x1, r.reshape(-1))
    c2 = ref.encrypt_batch(P, Q, x2, r.reshape(-1))
    t0 = time.perf_counter()
    ref.multpoly_batch(c1, D1, c2, D2, count)
    el = time.perf_counter() - t0
    return count * D1 * D2 / el, el


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    count = 16 * cores  # EMults per step (~0.25 s of one core each at 512 bit): every thread gets 16
    t0 = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        cpu_emults(max(1, cores // 4), cores)
    tot, el = 0.0, 0.0
    for _ in range(args.steps):
        v, e = cpu_emults(count, cores)
        tot += v * e
        el += e
    value = tot / el
    sample = ("%d steps x %d EMults (121 full pairings each) of the workload, C port of the oracle "
              "(oracle/cpu_ref.c), %d threads" % (args.steps, count, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64-limb integer (F_p, 519-bit)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "key_bits": KEY_BITS, "d1": D1, "d2": D2,
                                        "sample_emults_per_step": count},
        "cpu_baseline": {"value": value, "unit": "pairings/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
        "note": "the real reference (Go + cgo + libpbc + GMP) cannot be built on this box; this is a port",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms",
                                       "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from bgn_b200 import Engine, bench_imad_peak, workmodel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created; stdout must carry the
        # one JSON line only, so that banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    g = load_key()
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=local)
    L, EB, SB = eng.limbs, eng.elem_bytes, eng.scalar_bytes
    pairs = args.pairs

    # ---- synthetic inputs, built on the device (untimed): two batches of `pairs` x 11 level-1
    # coefficient ciphertexts of balanced base-3 digits with fresh randomness r < 2^(8*SB-1) <= n
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def make_batch(d):
        digits = torch.randint(-1, 2, (pairs * d,), generator=gen, device=dev, dtype=torch.int64)
        r = torch.randint(0, 256, (pairs * d, SB), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        return eng.encrypt_batch(digits, r.reshape(-1))

    c1, c2 = make_batch(D1), make_batch(D2)
    out = torch.empty(pairs * (D1 + D2) * EB, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- integer-pipe peak, measured live (IMAD.WIDE.U32 microkernel, 8 independent chains/thread)
    ms_peak, ipt = bench_imad_peak(local, 4096, 148 * 8, 256)
    imad_peak = 148 * 8 * 256 * ipt / (ms_peak * 1e-3)  # IMAD.WIDE instructions / s

    eng.timing_enable(True)

    def step_dev():
        flush.fill_(1)
        torch.cuda.synchronize()
        eng.multpoly_batch(c1, D1, c2, D2, pairs, out=out)
        return eng.timing_last_call()

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    eng.timing_reset()
    t_wall = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step_dev()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    k_ms, k_launches = eng.timing_get("k_miller")
    _, launches = eng.timing_get("")
    dev_ms = max_over_ranks(dev_ms)
    total_pairs = sum_over_ranks(float(pairs)) * args.steps
    value = total_pairs * D1 * D2 / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel (k_miller): executed 32x32->64 products / s vs the pipe
    modmuls_unit = workmodel.miller_unit_modmuls(p, n, l, D1, D2)  # F_p products incl. the lazily reduced ones
    prod_launch = pairs * workmodel.miller_unit_products(p, n, l, D1, D2)
    k_avg_s = (k_ms / max(1, k_launches)) * 1e-3
    achieved = prod_launch / k_avg_s
    algo_bytes = pairs * ((D1 + D2) * 2 * L * 4 + (D1 + D2 - 1) * 2 * L * 4)  # SoA in + out of k_miller
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "imad",
        "bound_note": "neither hbm nor tensor: north_star names the 32-bit integer multiply pipe (IMAD) as this path's "
                      "roofline; 50 000 products per byte moved -- the hbm object below shows memory at 0.002 % of its peak",
        "kernel": "k_miller<17>", "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
        "unit": "T(32x32->64 products)/s", "frac": achieved / imad_peak,
        "traffic": NCU_DRAM_BYTES_PER_UNIT * pairs,
        "traffic_note": "DRAM bytes per launch scaled from the ncu capture of one full wave (profiles/); algorithmic "
                        "bytes per launch = %d" % algo_bytes,
        "peak_source": "IMAD.WIDE.U32 microkernel measured in this run (nominal 148 SM x 64/clk x %.3f GHz = %.2f)" % (
            (clocks.get("sm_max_mhz") or 1965.0) / 1e3, 148 * 64 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12),
        "fp_products_per_emult": modmuls_unit, "products_per_emult": workmodel.miller_unit_products(p, n, l, D1, D2),
        "products_per_fused_modmul": workmodel.products_per_modmul(L),
        "kernel_ms": k_ms / max(1, k_launches), "kernel_share_of_step": k_ms / (dev_ms if world == 1 else max(dev_ms, 1e-9)),
        "hbm": {"algorithmic_GBs": algo_bytes / k_avg_s / 1e9, "peak_GBs": hbm_peak,
                "frac": algo_bytes / k_avg_s / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- end to end: pinned host buffers through the same C-ABI call
    h1 = torch.empty_like(c1, device="cpu").pin_memory()
    h2 = torch.empty_like(c2, device="cpu").pin_memory()
    ho = torch.empty_like(out, device="cpu").pin_memory()
    h1.copy_(c1)
    h2.copy_(c2)
    torch.cuda.synchronize()

    def step_e2e():
        flush.fill_(1)
        torch.cuda.synchronize()
        eng.multpoly_batch(h1, D1, h2, D2, pairs, out=ho)
        return eng.timing_last_call()

    step_e2e()
    barrier()
    e2e_ms = 0.0
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_ms += step_e2e()
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = sum_over_ranks(float(pairs)) * e2e_steps * D1 * D2 / (e2e_ms * 1e-3)
    same = bool((ho.to(dev) == out).all().item())  # host path and device path give identical bytes

    # ---- the other operations north_star names, measured briefly on the same key (device-resident
    # inputs, best of 2 calls after one warm-up; tools/opsbench.py has the per-kernel breakdown)
    def best_ms(fn, reps=2):
        fn()
        best = None
        for _ in range(reps):
            fn()
            t = eng.timing_last_call()
            best = t if best is None else min(best, t)
        return best

    n_enc = 1 << 18
    digits = torch.randint(-1, 2, (n_enc,), generator=gen, device=dev, dtype=torch.int64)
    rr = torch.randint(0, 256, (n_enc, SB), generator=gen, device=dev, dtype=torch.uint8)
    rr[:, 0] &= 0x3F
    rr = rr.reshape(-1)
    enc_out = torch.empty(n_enc * EB, dtype=torch.uint8, device=dev)
    enc_ms = best_ms(lambda: eng.encrypt_batch(digits, rr, out=enc_out))
    add_out = torch.empty((n_enc // 2) * EB, dtype=torch.uint8, device=dev)
    add_ms = best_ms(lambda: eng.g1_add_batch(enc_out[: (n_enc // 2) * EB], enc_out[(n_enc // 2) * EB:], out=add_out))
    eng.set_secret(int(g["q1"], 16), 1 << 20)
    n_dec = 1 << 14
    l2 = out[: n_dec * EB]  # level-2 coefficient ciphertexts produced by the timed EMult batch
    dec = {}

    def do_dec():
        dec["v"], dec["s"] = eng.decrypt_batch(l2, True)

    dec_ms = best_ms(do_dec)
    dec1 = {}

    def do_dec1():
        dec1["v"], dec1["s"] = eng.decrypt_batch(enc_out[: n_dec * EB], False)

    dec1_ms = best_ms(do_dec1)
    bl_out = torch.empty(n_dec * EB, dtype=torch.uint8, device=dev)
    bl1_ms = best_ms(lambda: eng.g1_blind_batch(enc_out, rr, out=enc_out.new_empty(n_enc * EB)))
    bl2_ms = best_ms(lambda: eng.gt_blind_batch(l2, rr[: n_dec * SB], out=bl_out))
    # the HBM-scale variant: 24-bit windows of Q (50 GB table per GPU, ~2 s to build, untimed)
    eng.set_option("enc_window", 24)
    enc24_ms = best_ms(lambda: eng.encrypt_batch(digits, rr, out=enc_out))
    eng.set_option("enc_window", 16)
    ops = {"encrypt_coeff_per_s": sum_over_ranks(n_enc / (enc_ms * 1e-3)),
           "encrypt_coeff_per_s_window24": sum_over_ranks(n_enc / (enc24_ms * 1e-3)),
           "eadd_coeff_per_s": sum_over_ranks((n_enc // 2) / (add_ms * 1e-3)),
           "decrypt_l2_per_s": sum_over_ranks(n_dec / (dec_ms * 1e-3)),
           "decrypt_l1_per_s": sum_over_ranks(n_dec / (dec1_ms * 1e-3)),
           "decrypt_l1_matches_plaintext": bool((dec1["v"] == digits[:n_dec]).all().item())
           and not bool(dec1["s"].any().item()),
           "rerandomize_l1_per_s": sum_over_ranks(n_enc / (bl1_ms * 1e-3)),
           "rerandomize_l2_per_s": sum_over_ranks(n_dec / (bl2_ms * 1e-3)),
           "decrypt_all_found": not bool(dec["s"].any().item()),
           "note": "keyBits=512, per-call device time incl. (de)serialisation kernels; Encrypt: x in {-1,0,1}, "
                   "512-bit r (16-bit windows of Q unless stated); Decrypt: trace of C^q1 by a Lucas ladder + one table probe over T=2^20 (level 1: pairing with P first); summed over ranks"}

    line = {
        "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32-limb integer (F_p, %d-bit, L=%d)" % (p.bit_length(), L),
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "key_bits": KEY_BITS, "pairs_per_gpu": pairs, "d1": D1, "d2": D2,
                   "parallelism": "independent pairs per GPU, no collective" if world > 1 else "1 GPU",
                   "l2": "256 MB flush between steps; step working set ~%d MB" % (
                       (pairs * (2 * (D1 + D2) * (EB + 2 * L * 4))) >> 20)},
        "emult_per_s": value / (D1 * D2),
        "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": int(h1.numel() + h2.numel()),
                "d2h_bytes_per_step": int(ho.numel()), "steps": e2e_steps, "bytes_match_device_path": same},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "wall_s": t_wall, "ops": ops,
    }
    if world == 1 and rank == 0 and not args.no_cpu:
        cores = host_cores()
        cnt = 48 * cores  # ~12 s on every core (~0.25 core-seconds per EMult at 512 bit)
        v, el = cpu_emults(cnt, cores)
        line["cpu_baseline"] = {
            "value": v, "unit": "pairings/s", "cores": cores, "kind": "port",
            "sample": "%d EMults (121 full pairings each) of the same workload in %.1f s, C port of the oracle "
                      "(oracle/cpu_ref.c), %d threads" % (cnt, el, cores)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 14)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
