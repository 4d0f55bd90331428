#!/bin/bash
# round 2, final call: the library at the end of the round (+ parabola steps, level-1 Decrypt as one pairing) -- parity tests, smoke,
# bench.py (driver-style), reference arm, opsbench, launch list, latencies, ncu --set full summaries
# (the .ncu-rep files are summarised on the box and deleted: gpurun_out/ brings back 64 MiB at most)
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r3t_pytest.log 2>&1
grep -E "passed|failed" $O/r3t_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/r3t_smoke.log 2>&1
tail -2 $O/r3t_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r3t_bench.json 2> $O/r3t_bench.err
tail -3 $O/r3t_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r3t_bench_ref.json 2>> $O/r3t_bench.err
timeout 600 python tools/opsbench.py > $O/r3t_ops.json 2> $O/r3t_ops.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r3t_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-verify --inner-length 4736 > $O/r3t_launches.log 2>&1
timeout 300 python tools/latency.py > $O/r3t_latency.json 2> $O/r3t_latency.err
for t in "miller k_miller" "fixed_pair k_miller_fixed_pair" "split k_miller_split" "miller1024 k_miller" "encrypt k_encrypt"; do
  set -- $t
  skip=0; [ "$1" = encrypt ] && skip=1
  timeout 600 ncu --set full --clock-control none -k regex:"^$2" -s $skip -c 1 -f -o $O/tmp_ncu_$1 python tools/ncu_targets.py $1 > $O/r3t_ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/tmp_ncu_$1.ncu-rep $O/r3t_ncu_$1.txt "$2 ($1), tools/ncu_targets.py $1" > /dev/null 2>> $O/r3t_ncu_$1.log
  rm -f $O/tmp_ncu_$1.ncu-rep
  grep -E "gpu__time_duration|fmaheavy|dram__bytes" $O/r3t_ncu_$1.txt | head -4
done
python - <<PY
import json
d=json.loads(open("$O/r3t_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"], d["roofline"]["traffic"])
print("strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")})
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
print("cpu", d.get("cpu_baseline"))
print(open("$O/r3t_bench_ref.json").read()[:200])
print(open("$O/r3t_latency.json").read())
PY
du -sh $O | tail -1
timeout 600 python tools/split_ab.py > $O/r3t_split_ab.json 2> $O/r3t_split_ab.err; tail -14 $O/r3t_split_ab.err | cut -c1-150
