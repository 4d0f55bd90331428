#!/bin/bash
# fused doubling-and-addition step: parity tests, bench, split A/B, 1024-bit wave
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r3h_pytest.log 2>&1
grep -E "passed|failed" $O/r3h_pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $O/r3h_bench.json 2> $O/r3h_bench.err
tail -3 $O/r3h_bench.err
python - <<PY
import json
d=json.loads(open("$O/r3h_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"], d["roofline"]["products_per_emult"])
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"], ip["ms_max_over_ranks"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
PY
timeout 600 python tools/split_ab.py > $O/r3h_split_ab.json 2> $O/r3h_split_ab.err; tail -14 $O/r3h_split_ab.err | cut -c1-150
