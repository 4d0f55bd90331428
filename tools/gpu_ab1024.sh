#!/bin/bash
# A/B of the 1024-bit MultPoly (config 5 shape, d = 8) for the default library and variants/lib*.so
O=gpurun_out; mkdir -p $O
N=${1:-4736}
run() { BGN_B200_LIB=$2 timeout 600 python tools/opsbench.py --key-bits 1024 --plaintexts 12000 --decrypts 64 --emults $N --d 8 --reps 2 2>>$O/ab1024_err.txt | python -c "
import json,sys
d=json.load(sys.stdin); e=d['ops']['emult_d8']
print('$1', 'emult/s=%.0f'%e['per_s'], 'frac=%.3f'%e['imad_frac'], 'k_miller_ms=%.1f'%e['kernel_ms']['k_miller'])"; }
run default ""
# variants/ and build/ are gpurun-ignored: copy the variant libraries to tools/_ab/
for f in tools/_ab/lib*.so; do [ -f "$f" ] && run $(basename $f .so) $PWD/$f; done
