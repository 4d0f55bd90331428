#!/bin/bash
# round 2, call B: parity tests, A/B of the two-warp general pairing, issue-mix microbenchmark (fixed), latency
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2b_pytest.log 2>&1
tail -15 $O/r2b_pytest.log
timeout 300 python tools/issuemix.py > $O/r2b_issuemix.json 2> $O/r2b_issuemix.err
tail -3 $O/r2b_issuemix.err
timeout 600 python tools/mapping_ab.py --which general --key-bits 512 > $O/r2b_pairduo_512.json 2> $O/r2b_ab.err
timeout 600 python tools/mapping_ab.py --which general --key-bits 1024 --max-log2 15 > $O/r2b_pairduo_1024.json 2>> $O/r2b_ab.err
timeout 600 python tools/mapping_ab.py --which fixed --key-bits 512 > $O/r2b_fixedpair_512.json 2>> $O/r2b_ab.err
tail -3 $O/r2b_ab.err
timeout 300 python tools/latency.py > $O/r2b_latency.json 2> $O/r2b_latency.err
cat $O/r2b_latency.json | head -40
python - <<PY
import json
for name in ("pairduo_512", "pairduo_1024", "fixedpair_512"):
    try:
        d=json.load(open("$O/r2b_%s.json" % name))
    except Exception as e:
        print(name, "missing", e); continue
    other = [k for k in d["products_per_pairing"] if k != "one_thread"][0]
    for r in d["sizes"]:
        print(name, r["count"], "1thr %.3f ms f=%.3f | %s %.3f ms f=%.3f | x%.2f eq=%s" % (r["one_thread"]["kernel_ms"], r["one_thread"]["imad_frac"], other, r[other]["kernel_ms"], r[other]["imad_frac"], r["speedup"], r["bytes_equal"]))
d=json.load(open("$O/r2b_issuemix.json"))
for k,v in d["mixes"].items():
    print(k, v["name"], "%.3f ms" % v["ms"], {a:round(b,2) for a,b in v.items() if isinstance(b,float) and a!="ms"})
PY
