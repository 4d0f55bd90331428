#!/bin/bash
# round 2, call Y: window test with r >= n, launch list of the committed library
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "windows or even_order or both_routes" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2y_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-verify --inner-length 4736 > $O/r2y_launches.log 2>&1
tail -1 $O/r2y_launches.log | cut -c1-200
grep -c . $O/r2y_launches.csv
