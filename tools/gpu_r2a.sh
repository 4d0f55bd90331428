#!/bin/bash
# round 2, call A: parity tests, lane-pair A/B of the fixed-argument pairing, issue-mix microbenchmark
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2a_pytest.log 2>&1
tail -15 $O/r2a_pytest.log
timeout 300 python tools/issuemix.py > $O/r2a_issuemix.json 2> $O/r2a_issuemix.err
tail -3 $O/r2a_issuemix.err
timeout 600 python tools/mapping_ab.py --key-bits 512 > $O/r2a_fixedpair_512.json 2> $O/r2a_fixedpair.err
timeout 600 python tools/mapping_ab.py --key-bits 1024 --max-log2 15 > $O/r2a_fixedpair_1024.json 2>> $O/r2a_fixedpair.err
tail -3 $O/r2a_fixedpair.err
python - <<PY
import json
for kb in (512, 1024):
    try:
        d=json.load(open("$O/r2a_fixedpair_%d.json" % kb))
    except Exception as e:
        print(kb, "missing", e); continue
    for r in d["sizes"]:
        print(kb, r["count"], "1thr %.3f ms f=%.3f | pair %.3f ms f=%.3f | x%.2f eq=%s" % (r["one_thread"]["kernel_ms"], r["one_thread"]["imad_frac"], r["lane_pair"]["kernel_ms"], r["lane_pair"]["imad_frac"], r["speedup_lane_pair"], r["bytes_equal"]))
d=json.load(open("$O/r2a_issuemix.json"))
for k,v in d["mixes"].items():
    print(k, v["name"], "%.3f ms" % v["ms"], {a:round(b,2) for a,b in v.items() if isinstance(b,float) and a!="ms"})
PY
