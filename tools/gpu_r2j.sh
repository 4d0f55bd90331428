#!/bin/bash
# round 2, call J: parity tests, wide team kernel A/B, bench.py
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2j_pytest.log 2>&1
tail -4 $O/r2j_pytest.log
timeout 900 python tools/wide_ab.py > $O/r2j_wide_ab.json 2> $O/r2j_wide_ab.err
cat $O/r2j_wide_ab.err | cut -c1-200
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-inner ) > $O/r2j_bench.json 2> $O/r2j_bench.err
tail -3 $O/r2j_bench.err
python - <<PY
import json
d=json.loads(open("$O/r2j_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")}, d["roofline"]["frac"])
PY
