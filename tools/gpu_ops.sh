#!/bin/bash
# Parity tests + per-operation benchmark.  Usage: tools/gpu_ops.sh <tag> [opsbench args]
TAG=${1:-ops}; shift
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
tail -4 $O/${TAG}_pytest.log
timeout 600 python tools/opsbench.py "$@" > $O/${TAG}_ops.json 2> $O/${TAG}_ops.err
tail -3 $O/${TAG}_ops.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_ops.json"))
for k,v in d["ops"].items():
    print("%-22s %12.0f /s  %8.3f ms  frac=%s  %s" % (k, v["per_s"], v["ms"], v.get("imad_frac"), {a:round(b,3) for a,b in v["kernel_ms"].items()}))
PY
