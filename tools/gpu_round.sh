#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, launch list, ncu capture of k_miller.  Usage: tools/gpu_round.sh <tag>
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
tail -5 $O/${TAG}_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/${TAG}_smoke.log 2>&1
tail -3 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json
timeout 300 python tools/primbench.py 1,10,20,11,12,13 > $O/${TAG}_primbench.txt 2>&1
cat $O/${TAG}_primbench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_miller -c 1 -f -o $O/${TAG}_miller python bench.py --steps 1 --warmup 0 --pairs 3404 --no-cpu > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
