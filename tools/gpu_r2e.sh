#!/bin/bash
# round 2, call E (2 GPUs): parity tests incl. handles and the C pthread driver on both devices; bench.py
# under torchrun at N=2 (strong scaling + config 5 with the NCCL all-gather) and in single-process mode
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2e_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2e_pytest.log 2>&1
tail -6 $O/r2e_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > $O/r2e_bench_n2.json 2> $O/r2e_bench_n2.err
tail -3 $O/r2e_bench_n2.err
( time timeout 600 python bench.py --single-process --gpus 2 --steps 3 --warmup 3 ) > $O/r2e_bench_sp2.json 2> $O/r2e_bench_sp2.err
tail -3 $O/r2e_bench_sp2.err

python - <<PY
import json
d=json.loads(open("$O/r2e_bench_n2.json").read().strip().splitlines()[0])
print({k: d[k] for k in ("value","ms_per_step","n_gpus","verified_units","verified_ok","gpu_launches")})
print("strong", {k:v for k,v in d["strong"].items() if k!="note"})
ip=d["inner_product"]; print("ip", {k:v for k,v in ip.items() if k not in ("config",)})
for k,v in d["ops"].items():
    if isinstance(v, dict): print(k, v["per_s"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel_ms"])
for f in ("r2e_bench_sp2.json","r2e_bench_sp1.json"):
    print(f, open("$O/"+f).read().strip()[:600])
PY
