#!/bin/bash
# parabola step also in the interleaved layout (1024 bit): parity tests at 1024 + all, config 5 timing
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r3d_pytest.log 2>&1
grep -E "passed|failed" $O/r3d_pytest.log
python tools/ip_timing.py 9472 2>&1 | grep call | cut -c1-200
