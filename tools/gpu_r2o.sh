#!/bin/bash
# round 2, call O: the final library -- parity tests, smoke, bench.py (driver-style), reference arm, launch list, opsbench
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2o_pytest.log 2>&1
grep -E "passed|failed" $O/r2o_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/r2o_smoke.log 2>&1
tail -2 $O/r2o_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r2o_bench.json 2> $O/r2o_bench.err
tail -3 $O/r2o_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2o_bench_ref.json 2>> $O/r2o_bench.err
timeout 600 python tools/opsbench.py > $O/r2o_ops.json 2> $O/r2o_ops.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2o_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-verify --inner-length 4736 > $O/r2o_launches.log 2>&1
timeout 300 python tools/latency.py > $O/r2o_latency.json 2> $O/r2o_latency.err
python - <<PY
import json
d=json.loads(open("$O/r2o_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"], d["roofline"]["traffic"])
print("strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")})
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
print("cpu", d.get("cpu_baseline"))
print(open("$O/r2o_bench_ref.json").read()[:200])
print(open("$O/r2o_latency.json").read())
PY
