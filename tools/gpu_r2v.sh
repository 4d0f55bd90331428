#!/bin/bash
# round 2, call V (N GPUs, N = $1): bench.py under torchrun with the final library (weak + strong scaling, configs 2 / 4
# per rank, config 5 with the NCCL all-gather)
N=${1:-8}
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 5 --warmup 3 ) > $O/r3v_bench_n$N.json 2> $O/r3v_bench_n$N.err
tail -2 $O/r3v_bench_n$N.err
python - <<PY
import json
N=$N
d=json.loads(open("$O/r3v_bench_n%d.json" % N).read().strip().splitlines()[0])
print(N, {k: d[k] for k in ("value","ms_per_step","n_gpus","verified_ok")})
print("  strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")}, d["strong"]["roofline"]["kernel_ms"])
ip=d["inner_product"]; print("  ip", ip["emult_per_s"], ip["ms_max_over_ranks"], ip["exchange_bytes_total"], ip["decrypted_matches_plaintext"])
for k,v in d["ops"].items():
    if isinstance(v, dict): print("  ", k, v["per_s"], v["ms"])
PY
