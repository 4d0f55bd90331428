#!/bin/bash
# k_miller<33> build variants: phase A product loops of 4 / 8 (shipped) / 16 rows, or unrolled; two waves of 8 x 8 products
for v in base a2 a8 a0; do
  lib=$PWD/abv/lib33_$v.so; [ $v = base ] && lib=$PWD/bgn_b200/libbgn_b200.so
  echo -n "$v "; BGN_B200_LIB=$lib timeout 300 python tools/ip_timing.py 9472 2>&1 | grep '"call": 2' | cut -c1-110
done
