#!/bin/bash
O=gpurun_out; mkdir -p $O
for t in 0 32 64 128; do
  BGN_MILLER_TAIL=$t timeout 300 python bench.py --no-cpu --steps 2 --warmup 2 2>>$O/tail_err.txt | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('tail=$t pairings/s=%.0f ms_per_step=%.2f match=%s' % (d['value'], d['ms_per_step'], d['e2e']['bytes_match_device_path']))"
done
BGN_MILLER_TAIL=32 timeout 600 python -m pytest tests -m gpu -x -q -k "multpoly or full_size or properties" 2>&1 | tail -2
BGN_MILLER_TAIL=32 timeout 300 python tools/inner_product.py --key-bits 1024 --length 8192 --slots 8 2>/dev/null | tail -1 | cut -c1-400
BGN_MILLER_TAIL=0 timeout 300 python tools/inner_product.py --key-bits 1024 --length 8192 --slots 8 2>/dev/null | tail -1 | cut -c1-400
