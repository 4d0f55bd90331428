#!/bin/bash
# k_miller<17> build variants against the shipped library: dedicated squarings inside the fused phase-A routines,
# phase A on the 8-row / 4-row product loop (smaller code).  Headline step time, 2^14 pairs.
O=gpurun_out
mkdir -p $O
for v in base fsqr loopa4 loopa2; do
  lib=$PWD/abv/lib_$v.so; [ $v = base ] && lib=$PWD/bgn_b200/libbgn_b200.so
  BGN_B200_LIB=$lib timeout 300 python bench.py --steps 6 --warmup 3 --no-inner --no-cpu --no-verify > $O/r3ab_$v.json 2> $O/r3ab_$v.err
  python -c "
import json; d=json.loads(open('$O/r3ab_$v.json').read().strip().splitlines()[0]); print('$v', round(d['ms_per_step'],2), 'ms', round(d['value']))"
done
