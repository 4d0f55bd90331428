#!/bin/bash
# Quick GPU pass: parity tests, primitive microbenchmarks, bench (no CPU baseline).  Usage: tools/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
tail -4 $O/${TAG}_pytest.log
timeout 300 python tools/primbench.py > $O/${TAG}_primbench.txt 2>&1
cat $O/${TAG}_primbench.txt
timeout 600 python bench.py --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
BGN_MILLER_GROUPS=2 timeout 600 python bench.py --no-cpu --pairs 16280 > $O/${TAG}_bench_g2.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_g2.json
