#!/bin/bash
# EAdd: more threads with fewer elements per shared inversion?
mkdir -p gpurun_out
for cfg in "8 37888" "6 56832" "4 75776" "4 113664" "2 151552"; do
  set -- $cfg
  timeout 300 python tools/affadd_k.py $1 $2 2> gpurun_out/r3z_$1_$2.err > /dev/null
  head -1 gpurun_out/r3z_$1_$2.err | cut -c1-220
done
