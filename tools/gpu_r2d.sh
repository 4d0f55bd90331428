#!/bin/bash
# round 2, call D: parity tests, general-pairing A/B with the shipped knobs, first run of the new bench.py
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2d_pytest.log 2>&1
tail -6 $O/r2d_pytest.log
timeout 600 python tools/mapping_ab.py --which general --key-bits 512 > $O/r2d_pairduo_512.json 2> $O/r2d_ab.err
timeout 600 python tools/mapping_ab.py --which general --key-bits 1024 --max-log2 15 > $O/r2d_pairduo_1024.json 2>> $O/r2d_ab.err
tail -3 $O/r2d_ab.err
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $O/r2d_bench.json 2> $O/r2d_bench.err
tail -5 $O/r2d_bench.err
python - <<PY
import json
for name in ("pairduo_512", "pairduo_1024"):
    try:
        d=json.load(open("$O/r2d_%s.json" % name))
    except Exception as e:
        print(name, "missing", e); continue
    other = [k for k in d["products_per_pairing"] if k != "one_thread"][0]
    for r in d["sizes"]:
        print(name, r["count"], "1thr %.3f ms f=%.3f | %s %.3f ms f=%.3f | x%.2f eq=%s" % (r["one_thread"]["kernel_ms"], r["one_thread"]["imad_frac"], other, r[other]["kernel_ms"], r[other]["imad_frac"], r["speedup"], r["bytes_equal"]))
d=json.loads(open("$O/r2d_bench.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","verified_units","verified_ok","gpu_launches")})
print("e2e", d["e2e"])
print("roofline frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("strong", d["strong"])
print("ip", d["inner_product"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
    else: print(k, v)
print("cpu", d.get("cpu_baseline"))
PY
