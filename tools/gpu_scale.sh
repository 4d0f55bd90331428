#!/bin/bash
# bench.py on N GPUs of the box (gpurun --gpus N).  Usage: tools/gpu_scale.sh <N>
N=${1:-2}; O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > $O/scale_n$N.json 2> $O/scale_n$N.err
wc -l $O/scale_n$N.json; python -c "
import json; d=json.loads(open('$O/scale_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['ops'])"
