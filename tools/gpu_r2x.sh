#!/bin/bash
# round 2, call X: the final library once more (level-1 Decrypt as one pairing, y = 0 guard) -- parity tests, smoke,
# bench.py (driver-style), reference arm
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2x_pytest.log 2>&1
grep -E "passed|failed" $O/r2x_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/r2x_smoke.log 2>&1
tail -2 $O/r2x_smoke.log | head -1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r2x_bench.json 2> $O/r2x_bench.err
tail -3 $O/r2x_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2x_bench_ref.json 2>> $O/r2x_bench.err
python - <<PY
import json
d=json.loads(open("$O/r2x_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"]); print("roofline frac", d["roofline"]["frac"], d["roofline"]["traffic"])
ip=d["inner_product"]; print("ip", ip["emult_per_s"], ip["roofline"]["frac"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
PY
