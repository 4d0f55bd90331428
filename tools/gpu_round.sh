#!/bin/bash
# One GPU-box pass: parity tests, smoke, both bench arms, launch list, ncu captures of the dominant kernels.
# Usage: tools/gpu_round.sh <tag>
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
head -5 $O/${TAG}_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/${TAG}_smoke.log 2>&1
tail -3 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_miller -c 1 -f -o $O/${TAG}_miller python bench.py --steps 1 --warmup 0 --pairs 3404 --no-cpu > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dec_lucas|k_miller_fixed|k_gt_blind' -c 3 -f -o $O/${TAG}_ops python tools/opsbench.py --reps 1 > $O/${TAG}_ncu_ops.log 2>&1
tail -2 $O/${TAG}_ncu_ops.log
