#!/bin/bash
# parabola step in the team kernel: parity tests, bench
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r3a_pytest.log 2>&1
grep -E "passed|failed" $O/r3a_pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $O/r3a_bench.json 2> $O/r3a_bench.err
tail -3 $O/r3a_bench.err
python - <<PY
import json
d=json.loads(open("$O/r3a_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"], d["roofline"]["products_per_emult"])
print("strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")})
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
PY
timeout 600 python tools/split_ab.py > $O/r3a_split_ab.json 2> $O/r3a_split_ab.err; tail -14 $O/r3a_split_ab.err | cut -c1-250
