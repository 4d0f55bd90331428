#!/bin/bash
# round 2, call W: level-1 Decrypt as one pairing with the line table of q1*P -- parity tests, opsbench, latency
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2w_pytest.log 2>&1
grep -E "passed|failed|Error" $O/r2w_pytest.log | head -5
timeout 600 python tools/opsbench.py > $O/r2w_ops.json 2> $O/r2w_ops.err
python - <<PY
import json
dd=json.load(open("$O/r2w_ops.json"))
for k,v in dd["ops"].items(): print("%-22s %12.0f /s %8.3f ms frac=%s %s" % (k, v["per_s"], v["ms"], v.get("imad_frac"), {a: round(b, 3) for a, b in v["kernel_ms"].items()}))
PY
timeout 300 python tools/latency.py > $O/r2w_latency.json 2> $O/r2w_latency.err
cat $O/r2w_latency.json
