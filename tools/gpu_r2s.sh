#!/bin/bash
# round 2, call S: Encrypt with the tables in twisted Edwards form -- parity tests, window x form sweep
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2s_pytest.log 2>&1
grep -E "passed|failed|Error|error" $O/r2s_pytest.log | head -5
timeout 900 python tools/enc_windows.py > $O/r2s_enc_windows.json 2> $O/r2s_enc_windows.err; cat $O/r2s_enc_windows.err | cut -c1-360
timeout 600 python tools/opsbench.py > $O/r2s_ops.json 2> $O/r2s_ops.err
python - <<PY
import json
dd=json.load(open("$O/r2s_ops.json"))
for k,v in dd["ops"].items(): print("%-22s %12.0f /s %8.3f ms frac=%s" % (k, v["per_s"], v["ms"], v.get("imad_frac")))
PY
