#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/opsbench.py --reps 2 --enc-window 24 > $O/ops_w24.json 2>$O/ops_w24.err; tail -2 $O/ops_w24.err
timeout 600 python bench.py --no-cpu --steps 2 > $O/bench_w.json 2>$O/bench_w.err; tail -2 $O/bench_w.err
python - <<PY
import json
d=json.load(open("$O/ops_w24.json"))
for k in ("encrypt","blind_l1"):
    v=d["ops"][k]; print(k, v["per_s"], v["ms"], v.get("imad_frac"), v["kernel_ms"])
b=json.load(open("$O/bench_w.json")); print(b["value"], b["ops"])
PY
