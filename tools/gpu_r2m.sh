#!/bin/bash
# round 2, call M: compute-sanitizer memcheck / racecheck / synccheck over the round-2 kernels (small keys)
O=gpurun_out
mkdir -p $O
K='two_warps or lane_pair or split_team or wide_team or handles or nondeterministic_poly or fixed_pairing'
( timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "(128 or 64 or 256) and ($K)" ) > $O/r2m_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" $O/r2m_memcheck.log | tail -3
( timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "(128 or 64) and (two_warps or lane_pair or split_team or wide_team or handles)" ) > $O/r2m_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" $O/r2m_racecheck.log | tail -3
( timeout 1500 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -x -k "(128 or 64) and (two_warps or lane_pair or split_team or wide_team)" ) > $O/r2m_synccheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" $O/r2m_synccheck.log | tail -3
python -m pytest tests/test_gpu_parity.py -q --collect-only -k "(128 or 64 or 256) and ($K)" 2>/dev/null | tail -25
