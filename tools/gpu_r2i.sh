#!/bin/bash
# round 2, call I: parity tests, bench.py (driver-style), reference arm, launch list, ncu --set full summaries
# (the .ncu-rep files are summarised on the box and deleted: gpurun_out/ brings back 64 MiB at most)
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2i_pytest.log 2>&1
tail -4 $O/r2i_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r2i_bench.json 2> $O/r2i_bench.err
tail -3 $O/r2i_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2i_bench_ref.json 2>> $O/r2i_bench.err
timeout 600 python tools/opsbench.py > $O/r2i_ops.json 2> $O/r2i_ops.err
timeout 600 python tools/mapping_ab.py --which general --key-bits 512 --max-log2 15 > $O/r2i_pairduo_512.json 2> $O/r2i_ab.err
BGN_B200_LIB=$PWD/tools/_ab/lib_duo_nosqr.so timeout 600 python tools/mapping_ab.py --which general --key-bits 512 --max-log2 15 > $O/r2i_pairduo_512_nosqr.json 2>> $O/r2i_ab.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2i_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-verify --inner-length 4736 > $O/r2i_launches.log 2>&1
tail -2 $O/r2i_launches.log | cut -c1-300
for t in "miller k_miller" "fixed_pair k_miller_fixed_pair" "pair_duo k_pair_duo" "split k_miller_split" "miller1024 k_miller" "dec_lucas k_dec_lucas"; do
  set -- $t
  timeout 600 ncu --set full --clock-control none -k regex:"^$2" -c 1 -f -o $O/tmp_ncu_$1 python tools/ncu_targets.py $1 > $O/r2i_ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/tmp_ncu_$1.ncu-rep $O/r2i_ncu_$1.txt "$2 ($1), tools/ncu_targets.py $1" > /dev/null 2>> $O/r2i_ncu_$1.log
  rm -f $O/tmp_ncu_$1.ncu-rep
  grep -E "gpu__time_duration|fmaheavy|dram__bytes" $O/r2i_ncu_$1.txt | head -4
done
python - <<PY
import json
d=json.loads(open("$O/r2i_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"])
print("strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")})
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
print("cpu", d.get("cpu_baseline"))
for name in ("pairduo_512", "pairduo_512_nosqr"):
    dd=json.load(open("$O/r2i_%s.json" % name))
    for r in dd["sizes"]:
        if r["count"] in (1, 4096, 16384, 32768): print(name, r["count"], "1thr %.3f | duo %.3f ms f=%.3f" % (r["one_thread"]["kernel_ms"], r["two_warps"]["kernel_ms"], r["two_warps"]["imad_frac"]))
dd=json.load(open("$O/r2i_ops.json"))
for k,v in dd["ops"].items(): print("%-22s %12.0f /s %8.3f ms frac=%s" % (k, v["per_s"], v["ms"], v.get("imad_frac")))
PY
du -sh $O | tail -1
