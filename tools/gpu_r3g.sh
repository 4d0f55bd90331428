#!/bin/bash
# parabola entries in the recorded table of P: parity tests, fixed-argument mapping A/B, latencies
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r3g_pytest.log 2>&1
grep -E "passed|failed" $O/r3g_pytest.log
timeout 600 python tools/mapping_ab.py --which fixed --key-bits 512 > $O/r3g_fixedpair_512.json 2> $O/r3g_ab.err
timeout 600 python tools/mapping_ab.py --which fixed --key-bits 1024 --max-log2 15 > $O/r3g_fixedpair_1024.json 2>> $O/r3g_ab.err
python - <<PY
import json
for kb in (512, 1024):
    d=json.load(open("$O/r3g_fixedpair_%d.json" % kb))
    for r in d["sizes"]:
        print(kb, r["count"], "1thr %.3f ms f=%.3f | pair %.3f ms f=%.3f" % (r["one_thread"]["kernel_ms"], r["one_thread"]["imad_frac"], r["lane_pair"]["kernel_ms"], r["lane_pair"]["imad_frac"]))
PY
timeout 300 python tools/latency.py > $O/r3g_latency.json 2> $O/r3g_latency.err
python -c "
import json; d=json.load(open('$O/r3g_latency.json')); print(d['kb512'])"
