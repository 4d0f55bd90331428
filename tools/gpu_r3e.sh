#!/bin/bash
# parabola step in the split team kernel: parity tests, split A/B
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r3e_pytest.log 2>&1
grep -E "passed|failed" $O/r3e_pytest.log
timeout 600 python tools/split_ab.py > $O/r3e_split_ab.json 2> $O/r3e_split_ab.err; tail -14 $O/r3e_split_ab.err | cut -c1-200
