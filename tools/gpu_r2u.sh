#!/bin/bash
# round 2, call U: compute-sanitizer memcheck / racecheck / synccheck over the kernels changed after call M
# (Edwards Encrypt + table conversion, scaled lines in the team / split kernels, normalised line table, lane pair)
O=gpurun_out
mkdir -p $O
K='encrypt or blind or make_l2 or multpoly or decrypt_l1 or lane_pair or split_team or make_poly_l2 or fixed_pairing'
( timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "(kb64 or kb128 or 128 or 64) and ($K) and not windows" ) > $O/r2u_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" $O/r2u_memcheck.log | tail -3
( timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "(kb64 or 64) and (multpoly or split_team or lane_pair or make_l2)" ) > $O/r2u_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" $O/r2u_racecheck.log | tail -3
( timeout 1500 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -x -k "(kb64 or 64) and (multpoly or split_team or lane_pair)" ) > $O/r2u_synccheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" $O/r2u_synccheck.log | tail -3
