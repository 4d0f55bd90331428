#!/bin/bash
# round 2, call K (8 GPUs): bench.py under torchrun at N = 8 and N = 4 (weak + strong scaling, config 5 with the
# NCCL all-gather), and one process driving all 8 GPUs
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
for N in 8 4; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 ) > $O/r2k_bench_n$N.json 2> $O/r2k_bench_n$N.err
  tail -2 $O/r2k_bench_n$N.err
done
( time timeout 600 python bench.py --single-process --gpus 8 --steps 5 --warmup 3 ) > $O/r2k_bench_sp8.json 2> $O/r2k_bench_sp8.err
tail -2 $O/r2k_bench_sp8.err
python - <<PY
import json
for N in (8, 4):
    d=json.loads(open("$O/r2k_bench_n%d.json" % N).read().strip().splitlines()[0])
    print(N, {k: d[k] for k in ("value","ms_per_step","n_gpus","verified_ok")})
    print("  strong", {k:v for k,v in d["strong"].items() if k not in ("note","roofline")}, d["strong"]["roofline"]["kernel_ms"])
    ip=d["inner_product"]; print("  ip", ip["emult_per_s"], ip["ms_max_over_ranks"], ip["exchange_bytes_total"], ip["decrypted_matches_plaintext"])
print(open("$O/r2k_bench_sp8.json").read()[:400])
PY
