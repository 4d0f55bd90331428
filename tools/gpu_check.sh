#!/bin/bash
# Validation pass on a GPU box: parity tests, smoke, both bench arms, ops benchmark.  Usage: tools/gpu_check.sh <tag>
TAG=${1:-chk}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
tail -5 $O/${TAG}_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/${TAG}_smoke.log 2>&1
tail -3 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_ref.json
timeout 300 python tools/opsbench.py > $O/${TAG}_ops.json 2> $O/${TAG}_ops.err
tail -3 $O/${TAG}_ops.err
