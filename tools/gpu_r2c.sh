#!/bin/bash
# round 2, call C: knobs of k_pair_duo + first run of the new bench.py
O=gpurun_out
mkdir -p $O
timeout 900 python tools/duo_ab.py > $O/r2c_duo_ab.json 2> $O/r2c_duo_ab.err
cat $O/r2c_duo_ab.err | tail -20
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $O/r2c_bench.json 2> $O/r2c_bench.err
tail -5 $O/r2c_bench.err
python - <<PY
import json
d=json.loads(open("$O/r2c_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"])
print("roofline frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("strong", d["strong"])
print("ip", d["inner_product"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
    else: print(k, v)
print("cpu", d.get("cpu_baseline"))
PY
