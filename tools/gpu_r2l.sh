#!/bin/bash
# round 2, call L: primbench of the interleaved lazy line_mul (mode 83 against 80)
O=gpurun_out
mkdir -p $O
timeout 300 python tools/primbench.py 80,83,40 > $O/r2l_primbench.txt 2>&1
cat $O/r2l_primbench.txt
