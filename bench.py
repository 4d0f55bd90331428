#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its headline config.

  metric   : reference-equivalent Tate pairings/s (= EMult/s x d1*d2) at 512-bit keys
  workload : config 3 -- batched EMult (MultPoly, poly.go:123-156) of 2^14 pairs of level-1
             polynomial ciphertexts with d1 = d2 = 11 coefficient slots (121 pairings and 22
             output slots per EMult); one "step" = one bgn_multpoly_batch call over the batch.
  N > 1    : every rank (one process per GPU) runs its own 2^14 pairs -- independent units, no
             data-path collective -> "scaling": "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

Beside the headline (config 3, weak scaling) the same line carries, nested where the driver keeps them
(`config`, `roofline`): config 3 as worded (2^14 pairs in TOTAL split over the ranks: `strong`), configs 2
and 4 at their stated sizes with a roofline object per dominant kernel (`roofline.ops`), and config 5
(keyBits=1024 inner product of length 2^16 with the NCCL all-gather + fold inside the timed region:
`config.inner_product_config5`).  `verified_units` EMults of the TIMED output are recomputed by the CPU
oracle in the same run.

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
def load_measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def dram_bytes_per_miller_unit():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_miller<17> launch divided by its units, from
    the newest `ncu --set full` capture summarised under profiles/ (tools/ncu_summary.py writes the json)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)):
        if name.endswith("_miller_dram.json"):
            best = os.path.join(pdir, name)
    if best:
        try:
            with open(best) as f:
                d = json.load(f)
            return d["dram_bytes"] / d["units"], os.path.basename(best)
        except Exception:
            pass
    return NCU_DRAM_BYTES_PER_UNIT, "r01_miller_v24_ncu.txt"


def verify_emults(g, c1, c2, out, pairs, EB, sample=8, seed=11):
    """Bit-exact check of `sample` EMults taken from the TIMED batch against the CPU port of the oracle
    (oracle/cpu_ref.c) run here, on this box, on the very inputs the GPU multiplied: -> (checked, ok)."""
    import numpy as np
    from oracle.cpu_ref import CpuRef
    ref = CpuRef(int(g["p"], 16), int(g["n"], 16), g["l"], threads=min(sample, host_cores()))
    rng = np.random.default_rng(seed)
    idx = sorted(set([0, pairs - 1] + [int(x) for x in rng.integers(0, pairs, max(0, sample - 2))]))
    a = np.concatenate([c1[i * D1 * EB:(i + 1) * D1 * EB].cpu().numpy() for i in idx])
    b = np.concatenate([c2[i * D2 * EB:(i + 1) * D2 * EB].cpu().numpy() for i in idx])
    exp = np.asarray(ref.multpoly_batch(a, D1, b, D2, len(idx))).reshape(-1)
    got = np.concatenate([out[i * (D1 + D2) * EB:(i + 1) * (D1 + D2) * EB].cpu().numpy() for i in idx])
    return len(idx), bool(exp.tobytes() == got.tobytes())


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from bgn_b200 import Engine, bench_imad_peak, workmodel
    from bgn_b200.multi import inner_product, shard_range

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

    def all_true(x: bool) -> bool:
        return sum_over_ranks(0.0 if x else 1.0) == 0.0

    g = load_key()
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=local)
    L, EB, SB = eng.limbs, eng.elem_bytes, eng.scalar_bytes
    pairs = args.pairs

    # ---- synthetic inputs, built on the device (untimed): two batches of `pairs` x 11 level-1
    # coefficient ciphertexts of balanced base-3 digits with fresh randomness r < 2^(8*SB-1) <= n
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def rand_scalars(count, nbytes=SB):
        r = torch.randint(0, 256, (count, nbytes), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        return r.reshape(-1)

    def make_batch(d):
        digits = torch.randint(-1, 2, (pairs * d,), generator=gen, device=dev, dtype=torch.int64)
        return eng.encrypt_batch(digits, rand_scalars(pairs * d))

    c1, c2 = make_batch(D1), make_batch(D2)
    out = torch.empty(pairs * (D1 + D2) * EB, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- integer-pipe peak, measured live (IMAD.WIDE.U32 microkernel, 8 independent chains/thread)
    ms_peak, ipt = bench_imad_peak(local, 4096, 148 * 8, 256)
    imad_peak = 148 * 8 * 256 * ipt / (ms_peak * 1e-3)  # IMAD.WIDE instructions / s
    ppm = workmodel.products_per_modmul(L)

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

    # ---- the timed output is checked HERE, on this box, against the CPU port of the oracle
    verified_units, verified_ok = 0, True
    if rank == 0 and not args.no_verify:
        verified_units, verified_ok = verify_emults(g, c1, c2, out, pairs, EB, sample=args.verify)

    # ---- roofline of the dominant kernel (k_miller): executed 32x32->64 products / s vs the pipe
    modmuls_unit = workmodel.miller_unit_modmuls(p, n, l, D1, D2)  # F_p products of the schedule with plain 5-product lines
    prod_unit = workmodel.miller_unit_products(p, n, l, D1, D2)
    k_avg_s = (k_ms / max(1, k_launches)) * 1e-3
    achieved = pairs * prod_unit / k_avg_s
    algo_bytes = pairs * ((D1 + D2) * 2 * L * 4 + (D1 + D2 - 1) * 2 * L * 4)  # SoA in + out of k_miller
    peaks = load_measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    dram_per_unit, dram_src = dram_bytes_per_miller_unit()
    roofline = {
        "bound": "imad",
        "bound_note": "neither hbm nor tensor: north_star names the 32-bit integer multiply pipe (IMAD) as this path's "
                      "roofline; 50 000 products per byte moved -- the hbm object below shows memory at 0.002 % of its peak",
        "kernel": "k_miller<17>", "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
        "unit": "T(32x32->64 products)/s", "frac": achieved / imad_peak,
        "traffic": dram_per_unit * pairs,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per unit from the ncu --set full capture summarised in "
                        "profiles/%s, times the units of one launch; algorithmic bytes per launch = %d" % (dram_src, algo_bytes),
        "peak_source": "IMAD.WIDE.U32 microkernel measured in this run (nominal 148 SM x 32/clk x %.3f GHz = %.2f); the "
                       "issue-mix microbenchmark (profiles/r02_issuemix.json) shows no instruction mix that multiplies "
                       "faster on this pipe" % ((clocks.get("sm_max_mhz") or 1965.0) / 1e3,
                                                148 * 32 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12),
        "modmul_equivalents_per_emult": prod_unit / (2 * L * L + L), "plain_line_modmuls_per_emult": modmuls_unit,
        "products_per_emult": prod_unit,
        "products_per_fused_modmul": ppm,
        "kernel_ms": k_ms / max(1, k_launches), "kernel_share_of_step": k_ms / (dev_ms if world == 1 else max(dev_ms, 1e-9)),
        "hbm": {"algorithmic_GBs": algo_bytes / k_avg_s / 1e9, "peak_GBs": hbm_peak,
                "frac": algo_bytes / k_avg_s / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- end to end: pinned host buffers through the same C-ABI call, every step
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
    e2e_steps = args.steps
    for _ in range(e2e_steps):
        e2e_ms += step_e2e()
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = sum_over_ranks(float(pairs)) * e2e_steps * D1 * D2 / (e2e_ms * 1e-3)
    same = bool((ho.to(dev) == out).all().item())  # host path and device path give identical bytes

    def best_ms(fn, prefix=None, reps=2):
        """best device ms of `reps` calls after one warm-up -> (call ms, ms of the kernels named `prefix`)"""
        fn()
        best = None
        for _ in range(reps):
            eng.timing_reset()
            fn()
            t = eng.timing_last_call()
            k = eng.timing_get(prefix)[0] if prefix else 0.0
            if best is None or t < best[0]:
                best = (t, k)
        return best

    def op_roofline(kernel, kernel_ms, units, products_per_unit):
        ach = units * products_per_unit / (kernel_ms * 1e-3)
        return {"bound": "imad", "kernel": kernel, "kernel_ms": kernel_ms, "units": units,
                "products_per_unit": products_per_unit, "achieved": ach / 1e12, "peak": imad_peak / 1e12,
                "unit": "T(32x32->64 products)/s", "frac": ach / imad_peak}

    # ---- config 3 AS WORDED: 2^14 pairs in total, sharded over the ranks (strong scaling)
    total_strong = 1 << 14
    lo, hi = shard_range(total_strong, rank, world)
    mine = min(hi - lo, pairs)
    o_s = out[: mine * (D1 + D2) * EB]
    barrier()
    t_s, k_s = best_ms(lambda: eng.multpoly_batch(c1[: mine * D1 * EB], D1, c2[: mine * D2 * EB], D2, mine, out=o_s),
                       "k_miller", reps=3)
    t_s_max = max_over_ranks(t_s)
    strong = {"pairs_total": int(sum_over_ranks(float(mine))), "pairs_per_gpu": mine, "ms": t_s_max,
              "pairings_per_s": sum_over_ranks(float(mine)) * D1 * D2 / (t_s_max * 1e-3),
              "emult_per_s": sum_over_ranks(float(mine)) / (t_s_max * 1e-3),
              "roofline": op_roofline("k_miller<17>", max_over_ranks(k_s), mine, prod_unit),
              "note": "BASELINE config 3 as worded: 2^14 pairs in TOTAL split by pair index over the ranks "
                      "(bgn_b200.multi.shard_range), no collective; best of 3 calls, max over ranks. Below one wave "
                      "(3404 units per GPU) a rank's k_miller has fewer than two warps per scheduler"}

    # ---- BASELINE configs 2 and 4 at their stated sizes, each with its dominant kernel's roofline
    eng.set_secret(int(g["q1"], 16), 1 << 20)
    q1 = int(g["q1"], 16)
    n_pt = 1 << 16                      # config 2: 2^16 plaintexts x 11 balanced base-3 digits
    n_enc = n_pt * D1
    digits = torch.randint(-1, 2, (n_enc,), generator=gen, device=dev, dtype=torch.int64)
    rr = rand_scalars(n_enc)
    enc_out = torch.empty(n_enc * EB, dtype=torch.uint8, device=dev)
    enc_bits = workmodel.enc_window_auto(SB, L)  # what enc_window = 0 (the default) builds: 20 bits, 3.7 GB
    enc_ms, enc_k = best_ms(lambda: eng.encrypt_batch(digits, rr, out=enc_out), "k_encrypt")
    enc_ref = enc_out.clone()
    half = n_enc // 2                   # config 2: 2^15 pairwise AddPoly = 360 448 coefficient additions
    add_out = torch.empty(half * EB, dtype=torch.uint8, device=dev)
    add_ms, add_k = best_ms(lambda: eng.g1_add_batch(enc_out[: half * EB], enc_out[half * EB:], out=add_out), "k_g1_affadd")
    # the same additions on device-resident handles (bgn_buf): no (de)serialisation around the kernel
    hA = eng.import_batch(1, enc_out[: half * EB])
    hB = eng.import_batch(1, enc_out[half * EB:])
    hR = eng.g1_add_h(hA, hB)
    addh_ms, addh_k = best_ms(lambda: eng.g1_add_h(hA, hB, out=hR), "k_g1_affadd")
    addh_same = bool((hR.to_bytes(out=torch.empty_like(add_out)) == add_out).all().item())
    for h in (hA, hB, hR):
        h.free()
    n_dec = 1 << 14                     # config 4: 2^14 level-2 ciphertexts, T = 2^20, half negative, 1 % zeros
    av = torch.randint(1, 1 << 10, (n_dec,), generator=gen, device=dev, dtype=torch.int64)
    bv = torch.randint(-(1 << 10) + 1, 1 << 10, (n_dec,), generator=gen, device=dev, dtype=torch.int64)
    bv[::100] = 0
    ca = eng.encrypt_batch(av, rand_scalars(n_dec))
    cb = eng.encrypt_batch(bv, rand_scalars(n_dec))
    l2 = torch.empty(n_dec * EB, dtype=torch.uint8, device=dev)
    pair_ms, pair_k = best_ms(lambda: eng.pair_batch(ca, cb, out=l2), "k_pair_duo")
    pair_kernel, pair_prod = "k_pair_duo<17>", workmodel.pair_duo_products(p, n, l)
    if pair_k == 0.0:
        pair_ms, pair_k = best_ms(lambda: eng.pair_batch(ca, cb, out=l2), "k_miller")
        pair_kernel, pair_prod = "k_miller<17> (team of 1)", workmodel.miller_unit_products(p, n, l, 1, 1)
    dec = {}

    def do_dec():
        dec["v"], dec["s"] = eng.decrypt_batch(l2, True)

    dec_ms, dec_k = best_ms(do_dec, "k_dec_lucas")
    dec_ok = bool((dec["v"] == av * bv).all().item()) and not bool(dec["s"].any().item())
    dec1 = {}

    def do_dec1():
        dec1["v"], dec1["s"] = eng.decrypt_batch(enc_out[: n_dec * EB], False)

    dec1_ms, dec1_k = best_ms(do_dec1, "k_miller_fixed")
    dec1_ok = bool((dec1["v"] == digits[:n_dec]).all().item()) and not bool(dec1["s"].any().item())
    n_bl = 1 << 18
    bl1_ms, _ = best_ms(lambda: eng.g1_blind_batch(enc_out[: n_bl * EB], rr[: n_bl * SB], out=enc_out.new_empty(n_bl * EB)))
    bl_out = torch.empty(n_dec * EB, dtype=torch.uint8, device=dev)
    bl2_ms, _ = best_ms(lambda: eng.gt_blind_batch(l2, rr[: n_dec * SB], out=bl_out))
    # the HBM-scale variant: 24-bit windows of Q (50 GB table per GPU, ~2 s to build, untimed)
    eng.set_option("enc_window", 16)
    enc16_ms, enc16_k = best_ms(lambda: eng.encrypt_batch(digits, rr, out=enc_out), "k_encrypt")
    enc_same = bool((enc_out == enc_ref).all().item())
    eng.set_option("enc_window", 24)
    enc24_ms, enc24_k = best_ms(lambda: eng.encrypt_batch(digits, rr, out=enc_out), "k_encrypt")
    enc_same = enc_same and bool((enc_out == enc_ref).all().item())
    eng.set_option("enc_window", 0)
    del enc_ref
    fixed_pair_prod = workmodel.miller_fixed_pair_products(p, n, l)
    ops = {
        "encrypt": {"config": "BASELINE config 2: 2^16 plaintexts x 11 digits = 720 896 coefficient encryptions, %d-bit windows "
                              "of Q (the default: %.1f GB table)" % (enc_bits, workmodel.enc_table_bytes(SB, enc_bits, L) / 1e9),
                    "window_bits": enc_bits,
                    "per_s": sum_over_ranks(n_enc / (enc_ms * 1e-3)), "plaintexts_per_s": sum_over_ranks(n_pt / (enc_ms * 1e-3)),
                    "ms": max_over_ranks(enc_ms), "bytes_equal_all_windows": all_true(enc_same),
                    "roofline": op_roofline("k_encrypt<17>", enc_k, n_enc, workmodel.encrypt_products(n, SB, enc_bits, L))},
        "encrypt_window16": {"config": "the same with 16-bit windows (285 MB table)", "per_s": sum_over_ranks(n_enc / (enc16_ms * 1e-3)),
                             "ms": max_over_ranks(enc16_ms),
                             "roofline": op_roofline("k_encrypt<17>", enc16_k, n_enc, workmodel.encrypt_products(n, SB, 16, L))},
        "encrypt_window24": {"config": "the same with 24-bit windows (50 GB table)", "per_s": sum_over_ranks(n_enc / (enc24_ms * 1e-3)),
                             "ms": max_over_ranks(enc24_ms),
                             "roofline": op_roofline("k_encrypt<17>", enc24_k, n_enc, workmodel.encrypt_products(n, SB, 24, L))},
        "eadd": {"config": "BASELINE config 2: 2^15 pairwise AddPoly = 360 448 level-1 coefficient additions",
                 "per_s": sum_over_ranks(half / (add_ms * 1e-3)), "ms": max_over_ranks(add_ms),
                 "roofline": op_roofline("k_g1_affadd<17>", add_k, half, workmodel.affadd_products(L)),
                 "roofline_note": "5 products and 1 squaring per addition; the shared inversion runs on the ALU pipe (division-step "
                                  "GCD) and the (de)serialisation kernels around it are HBM-side: this op is not "
                                  "multiply-bound, the fraction says how far"},
        "eadd_handles": {"config": "the same 360 448 additions on device-resident handles (bgn_g1_add_h): the byte format, its "
                                   "Montgomery conversion and the curve check are paid once at import, not per operation",
                         "per_s": sum_over_ranks(half / (addh_ms * 1e-3)), "ms": max_over_ranks(addh_ms),
                         "bytes_equal_byte_path": all_true(addh_same),
                         "roofline": op_roofline("k_g1_affadd<17>", addh_k, half, workmodel.affadd_products(L))},
        "mult_pairs": {"config": "2^14 plain Mult (one pairing each, bgn.go:294-314)", "per_s": sum_over_ranks(n_dec / (pair_ms * 1e-3)),
                       "ms": max_over_ranks(pair_ms), "roofline": op_roofline(pair_kernel, pair_k, n_dec, pair_prod)},
        "decrypt_l2": {"config": "BASELINE config 4: 2^14 level-2 ciphertexts, T = 2^20, half negative, 1 % zeros",
                       "per_s": sum_over_ranks(n_dec / (dec_ms * 1e-3)), "ms": max_over_ranks(dec_ms),
                       "plaintexts_match": all_true(dec_ok),
                       "roofline": op_roofline("k_dec_lucas<17>", dec_k, n_dec, workmodel.dec_lucas_modmuls(q1) * ppm)},
        "decrypt_l1": {"config": "2^14 level-1 ciphertexts (pairing with P on a lane pair, then the ladder)",
                       "per_s": sum_over_ranks(n_dec / (dec1_ms * 1e-3)), "ms": max_over_ranks(dec1_ms),
                       "plaintexts_match": all_true(dec1_ok),
                       "roofline": op_roofline("k_miller_fixed_pair<17>", dec1_k, n_dec, fixed_pair_prod)},
        "rerandomize_l1_per_s": sum_over_ranks(n_bl / (bl1_ms * 1e-3)),
        "rerandomize_l2_per_s": sum_over_ranks(n_dec / (bl2_ms * 1e-3)),
        "note": "keyBits=512; per-call device time incl. (de)serialisation kernels, best of 2 after a warm-up, inputs "
                "resident in HBM; rates summed over ranks (every rank runs the stated size)"}
    del enc_out, add_out, digits, rr, flush

    # ---- BASELINE config 5: keyBits=1024 encrypted inner product, length 2^16, d = 8, sharded by index;
    # per-GPU GT product tree -> NCCL all-gather of the serialised partials -> fold kernel, all inside
    # the timed region; rank 0 decrypts the 16 result slots and compares with the plaintext result
    ip = None
    if not args.no_inner:
        ip = run_inner_product(args, rank, world, local, dev, barrier, max_over_ranks, sum_over_ranks)

    line = {
        "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32-limb integer (F_p, %d-bit, L=%d)" % (p.bit_length(), L),
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "key_bits": KEY_BITS, "pairs_per_gpu": pairs, "d1": D1, "d2": D2,
                   "parallelism": "independent pairs per GPU, no collective" if world > 1 else "1 GPU",
                   "l2": "256 MB flush between steps; step working set ~%d MB" % (
                       (pairs * (2 * (D1 + D2) * (EB + 2 * L * 4))) >> 20),
                   "strong_scaling_config3": strong, "inner_product_config5": ip},
        "emult_per_s": value / (D1 * D2),
        "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": int(h1.numel() + h2.numel()),
                "d2h_bytes_per_step": int(ho.numel()), "steps": e2e_steps, "bytes_match_device_path": same},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "wall_s": t_wall,
        "verified_units": verified_units, "verified_ok": verified_ok,
        "verified_note": "EMults sampled from the timed batch, recomputed on this box by the CPU port of the oracle "
                         "(121 full pairings each) and compared byte for byte",
    }
    roofline["ops"] = ops
    roofline["strong_scaling_config3"] = strong["roofline"]
    line["strong"] = strong
    line["inner_product"] = ip
    line["ops"] = ops
    if world == 1 and rank == 0 and not args.no_cpu:
        cores = host_cores()
        cnt = 48 * cores  # ~12 s on every core (~0.25 core-seconds per EMult at 512 bit)
        v, el = cpu_emults(cnt, cores)
        line["cpu_baseline"] = {
            "value": v, "unit": "pairings/s", "cores": cores, "kind": "port",
            "sample": "%d EMults (121 full pairings each) of the same workload in %.1f s, C port of the oracle "
                      "(oracle/cpu_ref.c), %d threads" % (cnt, el, cores)}
    if rank == 0:
        if not verified_ok:
            line["error"] = "timed output differs from the CPU oracle"
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and not verified_ok:
        sys.exit(3)


def run_inner_product(args, rank, world, local, dev, barrier, max_over_ranks, sum_over_ranks):
    import torch
    import torch.distributed as dist

    from bgn_b200 import Engine, workmodel
    from bgn_b200.multi import inner_product, shard_range
    kb, length, d = 1024, args.inner_length, 8
    with open(os.path.join(ROOT, "tests", "golden", "kb%d.json" % kb)) as f:
        g = json.load(f)
    p, n, l, q1 = int(g["p"], 16), int(g["n"], 16), g["l"], int(g["q1"], 16)
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=local)
    SB = eng.scalar_bytes
    lo, hi = shard_range(length, rank, world)
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
    cu = eng.encrypt_batch(u.reshape(-1), rnd())
    cv = eng.encrypt_batch(v.reshape(-1), rnd())
    enc_ms = eng.timing_last_call()
    EB = eng.elem_bytes
    warm = min(cnt, 256)
    inner_product(eng, cu[: warm * d * EB], d, cv[: warm * d * EB], d, warm)  # warm-up: module load, NCCL channel set-up
    barrier()
    eng.timing_reset()
    tm = {}
    t0 = time.perf_counter()
    total = inner_product(eng, cu, d, cv, d, cnt, timings=tm)
    torch.cuda.synchronize()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    k_ms = eng.timing_get("k_miller")[0]
    dev_ms = tm["multpoly_ms"] + tm["tree_ms"] + tm["fold_ms"]
    step_ms = max_over_ranks(dev_ms + tm["allgather_wall_ms"])
    conv = torch.zeros(2 * d, dtype=torch.int64, device=dev)
    for i in range(d):
        for k in range(d):
            conv[i + k] += (u[:, i] * v[:, k]).sum()
    if world > 1:
        dist.all_reduce(conv, op=dist.ReduceOp.SUM)
    ok, slots = True, []
    if rank == 0:
        eng.set_secret(q1, 1 << 20)
        vals, status = eng.decrypt_batch(total, True)
        slots = [int(x) for x in vals.cpu()]
        ok = (not bool(status.any().item())) and slots == [int(x) for x in conv.cpu()]
    prod_unit = workmodel.miller_unit_products(p, n, l, d, d)
    from bgn_b200 import bench_imad_peak
    ms_peak, ipt = bench_imad_peak(local, 2048, 148 * 8, 256)
    peak = 148 * 8 * 256 * ipt / (ms_peak * 1e-3)
    miller_ms = max_over_ranks(tm["multpoly_ms"])
    res = {
        "config": "BASELINE config 5: keyBits=1024 fixed-point encrypted inner product, length %d, d=%d slots, "
                  "sharded by index over %d GPU(s)" % (length, d, world),
        "emult_per_s": length / (step_ms * 1e-3), "pairings_per_s": length * d * d / (step_ms * 1e-3),
        "ms": step_ms, "wall_ms": max_over_ranks(wall_ms),
        "ms_max_over_ranks": {"multpoly": miller_ms, "l2_sum_tree": max_over_ranks(tm["tree_ms"]),
                              "allgather_wall": max_over_ranks(tm["allgather_wall_ms"]),
                              "fold": max_over_ranks(tm["fold_ms"]), "encrypt_untimed": max_over_ranks(enc_ms)},
        "collective": "NCCL all_gather_into_tensor of %d serialised GT elements per rank, device to device" % (2 * d)
                      if world > 1 else "none (1 GPU)",
        "exchange_bytes_per_rank": tm["exchange_bytes_per_rank"],
        "exchange_bytes_total": int(sum_over_ranks(float(tm["exchange_bytes_per_rank"]))) if world > 1 else 0,
        "decrypted_matches_plaintext": bool(ok), "slots": slots,
        "roofline": {"bound": "imad", "kernel": "k_miller<33>", "kernel_ms": max_over_ranks(k_ms), "units_per_gpu": cnt,
                     "products_per_unit": prod_unit, "achieved": cnt * prod_unit / (max_over_ranks(k_ms) * 1e-3) / 1e12,
                     "peak": peak / 1e12, "unit": "T(32x32->64 products)/s",
                     "frac": cnt * prod_unit / (max_over_ranks(k_ms) * 1e-3) / peak},
        "limit": "the Miller kernel: tree + all-gather + fold are %.2f %% of the step" % (
            100.0 * (step_ms - miller_ms) / step_ms),
    }
    eng.close()
    return res


def run_single_process(args):
    """ONE process driving `--gpus` contexts, one host thread each (the shape of a Go integration: one
    goroutine per device around the cgo calls).  Same workload and timing as the headline arm; the
    threads meet at a barrier before and after the timed steps and the slowest thread's device time
    counts.  Compare `value` with the torchrun arm's (one process per GPU) at the same N."""
    import threading

    import torch

    from bgn_b200 import Engine
    n = args.gpus
    g = load_key()
    p, nn, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    pairs = args.pairs
    res = [None] * n
    errs = []
    bar = threading.Barrier(n)

    def work(i):
        try:
            dev = torch.device("cuda", i)
            eng = Engine(p, nn, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=i)
            EB, SB = eng.elem_bytes, eng.scalar_bytes
            gen = torch.Generator(device=dev)
            gen.manual_seed(1234 + i)

            def make_batch(d):
                digits = torch.randint(-1, 2, (pairs * d,), generator=gen, device=dev, dtype=torch.int64)
                r = torch.randint(0, 256, (pairs * d, SB), generator=gen, device=dev, dtype=torch.uint8)
                r[:, 0] &= 0x3F
                return eng.encrypt_batch(digits, r.reshape(-1))

            c1, c2 = make_batch(D1), make_batch(D2)
            out = torch.empty(pairs * (D1 + D2) * EB, dtype=torch.uint8, device=dev)
            eng.timing_enable(True)
            for _ in range(args.warmup):
                eng.multpoly_batch(c1, D1, c2, D2, pairs, out=out)
            torch.cuda.synchronize(dev)
            bar.wait()
            t0 = time.perf_counter()
            ms = 0.0
            for _ in range(args.steps):
                eng.multpoly_batch(c1, D1, c2, D2, pairs, out=out)
                ms += eng.timing_last_call()
            torch.cuda.synchronize(dev)
            bar.wait()
            res[i] = (ms, time.perf_counter() - t0)
            eng.close()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
            try:
                bar.abort()
            except Exception:
                pass

    ths = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        print(json.dumps({"impl": "single-process", "error": errs}), flush=True)
        sys.exit(2)
    dev_ms = max(r[0] for r in res)
    wall = max(r[1] for r in res)
    value = n * pairs * args.steps * D1 * D2 / (dev_ms * 1e-3)
    # the same batch through ONE call of the multi-device handle (bgn_group_multpoly_batch: the library cuts the
    # batch into per-device shards and runs them from its own host threads), pinned host buffers, wall clock
    from bgn_b200 import EngineGroup
    grp = EngineGroup(p, nn, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), devices=list(range(n)))
    EB, SB = grp.elem_bytes, grp.scalar_bytes
    total = n * pairs
    rng = torch.Generator()
    rng.manual_seed(99)

    def host_batch(d):
        digits = torch.randint(-1, 2, (total * d,), generator=rng, dtype=torch.int64).pin_memory()
        r = torch.randint(0, 256, (total * d, SB), generator=rng, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        out = torch.empty(total * d * EB, dtype=torch.uint8).pin_memory()
        grp.encrypt_batch(digits, r.reshape(-1).pin_memory(), out=out)
        return out

    h1, h2 = host_batch(D1), host_batch(D2)
    ho = torch.empty(total * (D1 + D2) * EB, dtype=torch.uint8).pin_memory()
    grp.multpoly_batch(h1, D1, h2, D2, total, out=ho)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        grp.multpoly_batch(h1, D1, h2, D2, total, out=ho)
    group_wall = time.perf_counter() - t0
    grp.close()
    print(json.dumps({
        "impl": "single-process", "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "wall_s": wall,
        "value_by_wall_clock": n * pairs * args.steps * D1 * D2 / wall, "higher_is_better": True, "scaling": "weak",
        "group_call": {"value": total * args.steps * D1 * D2 / group_wall, "unit": "pairings/s", "ms_per_step": 1e3 * group_wall / args.steps,
                       "note": "ONE bgn_group_multpoly_batch call per step over all GPUs, pinned host buffers, wall clock "
                               "(host<->device copies inside)"},
        "config": {"workload": WORKLOAD, "pairs_per_gpu": pairs, "parallelism": "one process, one host thread and one bgn_ctx per GPU, no collective"},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 14)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="one process, one host thread + context per GPU (instead of torchrun's process per GPU)")
    ap.add_argument("--no-inner", action="store_true", help="skip the keyBits=1024 inner product (config 5)")
    ap.add_argument("--inner-length", type=int, default=1 << 16)
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--verify", type=int, default=8, help="EMults of the timed batch re-computed by the CPU oracle")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
