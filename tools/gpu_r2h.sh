#!/bin/bash
# round 2, call H: bench.py (1 GPU, driver-style steps), reference arm, launch list, ncu --set full captures
O=gpurun_out
mkdir -p $O
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r2h_bench.json 2> $O/r2h_bench.err
tail -3 $O/r2h_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2h_bench_ref.json 2>> $O/r2h_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-verify --inner-length 4736 > $O/r2h_launches.log 2>&1
tail -2 $O/r2h_launches.log
for t in "miller k_miller" "fixed_pair k_miller_fixed_pair" "pair_duo k_pair_duo" "split k_miller_split" "miller1024 k_miller"; do
  set -- $t
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$2" -c 1 -f -o $O/r2h_ncu_$1 python tools/ncu_targets.py $1 > $O/r2h_ncu_$1.log 2>&1
  tail -1 $O/r2h_ncu_$1.log
done
python - <<PY
import json
d=json.loads(open("$O/r2h_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"])
print("strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")})
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
print("cpu", d.get("cpu_baseline"))
print(open("$O/r2h_bench_ref.json").read()[:300])
PY
