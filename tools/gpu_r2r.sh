#!/bin/bash
# round 2, call R: line table of P normalised by the third coefficient (MillerFixed::record, pairlane.cuh)
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2r_pytest.log 2>&1
grep -E "passed|failed" $O/r2r_pytest.log
timeout 600 python tools/mapping_ab.py --which fixed --key-bits 512 > $O/r2r_fixedpair_512.json 2> $O/r2r_ab.err
timeout 600 python tools/mapping_ab.py --which fixed --key-bits 1024 --max-log2 15 > $O/r2r_fixedpair_1024.json 2>> $O/r2r_ab.err
tail -40 $O/r2r_ab.err | cut -c1-300
timeout 300 python tools/latency.py > $O/r2r_latency.json 2> $O/r2r_latency.err
cat $O/r2r_latency.json
